"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

CPU restatement of the two CLIP-side video components of SURVEY.md 8f-3, plain fp32/fp64 PyTorch functional code:

* `clip_st_adapter`  -- SpatioTemporalConvAdapter.forward, model/llava/model/multimodal_encoder/modeling_clip.py:598-612
* `adaptive_avgpool3d_tokens` -- AdaptiveAvgPooling3D.forward, model/llava/model/multimodal_encoder/pooling.py:15-25

Pinned to the reference classes themselves (`oracle/make_golden.py` executes their source from /root/reference and writes
tests/golden/clip_adapters.npz; `tests/test_oracle_golden.py` compares)."""
import torch
import torch.nn.functional as F


def clip_st_adapter(x, weight, bias, alpha):
    """x [(b t), 1 + h*w, c] with t = 8, h = 16 (modeling_clip.py:603); weight [c, c, 3, 3, 3]; returns the same shape."""
    BT, L, C = x.shape
    cls_embed, seq = x[:, :1], x[:, 1:]                                     # :600-601
    b = BT // 8
    v = seq.reshape(b, 8, 16, (L - 1) // 16, C).permute(0, 4, 1, 2, 3)             # '(b t) (h w) c -> b c t h w', t=8, h=16  (:603)
    y = torch.tanh(alpha) * F.relu(F.conv3d(v, weight, bias, padding="same")) + v  # :605
    y = y.permute(0, 2, 3, 4, 1).reshape(BT, L - 1, C)                             # 'b c t h w -> (b t) (h w) c'  (:607)
    return torch.cat((cls_embed, y), dim=1)                                        # :609


def adaptive_avgpool3d_tokens(x, num_frames=8, out_hw=(8, 9)):
    """x [(b t), h*w, c] -> [b, t*8*9, c]  (pooling.py:18-24; the module pools to (num_frames, 8, 9), :13)"""
    BT, N, C = x.shape
    h = w = int(N ** 0.5)
    v = x.reshape(BT // num_frames, num_frames, h, w, C).permute(0, 4, 1, 2, 3)
    v = F.adaptive_avg_pool3d(v, (num_frames, out_hw[0], out_hw[1]))
    return v.permute(0, 2, 3, 4, 1).reshape(BT // num_frames, -1, C)

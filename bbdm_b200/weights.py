"""Derived conv weights (host-side, at cache-refresh time only).

``upsample_phase_weights``: nearest-2x upsampling followed by a 3x3 'same' conv
(reference openaimodel.py:118 + :207, i.e. ResBlock(up=True).in_layers, :212-214) equals, per
output phase (a, b) = (y % 2, x % 2), a 2x2 conv on the LOW-RES input whose taps are sums of the
3x3 taps:

    upsampled row 2i+a+dy comes from source row i + floor((a+dy)/2),  dy in {-1,0,1}
      a = 0:  source rows (i-1, i)   get  (w[-1],        w[0] + w[1])
      a = 1:  source rows (i,  i+1)  get  (w[-1] + w[0], w[1])
    (same along x).  Out-of-image source rows are zero in both formulations.

4 phases x 4 taps = 16 [Cout, Cin] matrices -> 2.25x fewer MACs than 9 taps on 4x the pixels.
"""
import torch


def upsample_phase_weights(w: torch.Tensor) -> torch.Tensor:
    """w [Cout, Cin, 3, 3] -> [Cout, Cin, 16], index phase*4 + r*2 + c with phase = a*2 + b."""
    assert w.dim() == 4 and w.shape[2] == 3 and w.shape[3] == 3
    # row-combination matrices R[a][r, dy]: which 3x3 rows feed source row r of phase a
    R = torch.tensor([[[1., 0., 0.], [0., 1., 1.]],      # a = 0: r=0 <- dy=-1 ; r=1 <- dy=0,+1
                      [[1., 1., 0.], [0., 0., 1.]]],     # a = 1: r=0 <- dy=-1,0 ; r=1 <- dy=+1
                     dtype=w.dtype, device=w.device)
    # out[o,i,a,b,r,c] = sum_{dy,dx} R[a,r,dy] * R[b,c,dx] * w[o,i,dy,dx]
    out = torch.einsum("ary,bcx,oiyx->oiabrc", R, R, w)
    return out.reshape(w.shape[0], w.shape[1], 16).contiguous()

"""GPU diagnostic: which TMEM lanes hold the rows of an M = 64 tcgen05 accumulator (csrc/gemm.cu reads it with 32x32b loads:
row r sits in lane 32*(r/16) + r%16).  C[r, :] = r + 1 for a [128, 64] x [64, 64]^T problem on the 64-row tile path."""
import sys, torch
sys.path.insert(0, ".")
from textboost_b200 import _cabi as C, ops
M, N, K = 128, 64, 64
a = torch.zeros(M, K, device="cuda", dtype=torch.float16)
a[:, 0] = torch.arange(1, M + 1, device="cuda").half()
w = torch.zeros(N, K, device="cuda", dtype=torch.float16)
w[:, 0] = 1
out = ops.gemm(a, w, out_kind=C.TB_OUT_F32)
torch.cuda.synchronize()
print("col0 of first 64 rows:", out[:64, 0].int().tolist())
print("col0 of rows 64..127:", out[64:, 0].int().tolist())
print("row 0 first 8 cols:", out[0, :8].tolist(), "row 5:", out[5, :8].tolist())

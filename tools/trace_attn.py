"""Debug: per-role clock64() time line of CTA 0 of the persistent tcgen05 attention kernels (vit_attention_tc.cu, TRF / TRB / TRI probes;
third item of the CTA, plus the start / end of every item).  Needs the probe build:
    make -C clip_diffusion_b200/csrc trace
    CLIPGUIDE_B200_LIB=$PWD/clip_diffusion_b200/csrc/libclipguide_b200_trace.so python tools/trace_attn.py 64 257 fwd|bwd
Output format: see the header lines of profiles/r02_attention_timeline_*.txt."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_diffusion_b200 import _lib
P = _lib.ptr
lib = _lib.load()
n, T, heads = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 257, 16
which = sys.argv[3] if len(sys.argv) > 3 else "fwd"
D = heads * 64
qkv = torch.randn(n * T, 3 * D, device="cuda").bfloat16(); ctx = torch.empty(n * T, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(n, heads, T, device="cuda"); dctx = torch.randn(n * T, D, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv); delta = torch.empty_like(lse)
def run():
    if which == "fwd":
        _lib.call("cg_attention_fwd", P(qkv), n, T, heads, P(ctx), P(lse))
    else:
        _lib.call("cg_attention_bwd", P(qkv), P(ctx), P(dctx), P(lse), n, T, heads, P(dqkv), P(delta))
_lib.call("cg_attention_fwd", P(qkv), n, T, heads, P(ctx), P(lse))
for _ in range(3):
    run()
trace = torch.zeros(16 * 64 * 8, dtype=torch.int64, device="cuda")
lib.cg_debug_attention_trace.argtypes = [ctypes.c_void_p]
lib.cg_debug_attention_trace(ctypes.c_void_p(trace.data_ptr()))
run()
torch.cuda.synchronize()
lib.cg_debug_attention_trace(None)
t = trace.cpu().view(16, 64, 8)
t0 = int(t[t > 0].min())
print("# %s n=%d T=%d: cycles relative to the first stamp; per role: block: events" % (which, n, T))
for role in range(16):
    rows = []
    for g in range(64):
        ev = t[role, g]
        if (ev > 0).any():
            rows.append("%d:[%s]" % (g, " ".join(str(int(e) - t0) if e > 0 else "-" for e in ev)))
    if rows:
        print("warp %2d  %s" % (role, "  ".join(rows)))

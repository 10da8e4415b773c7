"""GEMM tile-shape sweep on the ViT layer shapes of the bench workloads: every (M, N, K, epilogue) with the forced variants
(CG_GEMM_BN=128 / 256, CG_GEMM_PAIR=1 / 0 are read once per process => one subprocess per variant) and the automatic choice.
Usage (GPU box): python tools/bench_gemm_shapes.py > profiles/<name>.txt"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = []
for M in (16448, 8224, 4112, 2056, 6304, 3152, 1576, 788, 3200):  # L/14 x 64 / 32 / 16 / 8 cutouts, B/16 x 32 / 16 / 8 / 4, B/32 x 64
    D = 1024 if M in (16448, 8224, 4112, 2056) else 768
    SHAPES += [(M, 3 * D, D, "bias_bf16"), (M, D, D, "bias_resid_f32"), (M, 4 * D, D, "bias_qgelu"), (M, D, 4 * D, "bias_resid_f32"),
               (M, 4 * D, D, "dqgelu"), (M, D, 4 * D, "f32"), (M, D, D, "bf16"), (M, D, 3 * D, "f32")]

if len(sys.argv) > 1 and sys.argv[1] == "worker":
    import torch
    from clip_diffusion_b200 import _lib, vit_ops
    EPI = {"bias_bf16": _lib.EPI_BIAS_BF16, "bias_resid_f32": _lib.EPI_BIAS_RESID_F32, "bias_qgelu": _lib.EPI_BIAS_QGELU_BF16, "dqgelu": _lib.EPI_DQGELU_BF16,
           "f32": _lib.EPI_F32, "bf16": _lib.EPI_BF16}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for (M, N, K, en) in SHAPES:
        epi = EPI[en]
        a = torch.randn(M, K, device="cuda").bfloat16(); b = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16(); bias = torch.zeros(N, device="cuda")
        f32 = en in ("bias_resid_f32", "f32")
        out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
        aux = torch.zeros(M, N, device="cuda", dtype=torch.float32 if en == "bias_resid_f32" else torch.bfloat16)
        ts = []
        for it in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); vit_ops.gemm_bf16_tn(a, b, epi, bias=bias, out=out, aux=aux); e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        res["%d,%d,%d,%s" % (M, N, K, en)] = ts[len(ts) // 2]
    print(json.dumps(res))
    sys.exit(0)

variants = [("auto", {}), ("bn256", {"CG_GEMM_BN": "256", "CG_GEMM_PAIR": "0"}), ("bn128", {"CG_GEMM_BN": "128"}), ("pair", {"CG_GEMM_PAIR": "1"})]
results = {}
for name, env in variants:
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "worker"], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        print("# variant %s failed: %s" % (name, r.stderr[-500:]))
        continue
    results[name] = json.loads(r.stdout.strip().splitlines()[-1])
print("# us per GEMM (median of 7, L2 flushed); TFLOP/s of the automatic choice; best forced variant")
print("%-40s %9s %9s %9s %9s   %8s  %s" % ("M,N,K,epilogue", "auto", "bn256", "bn128", "pair", "auto TF/s", "best"))
tot = {k: 0.0 for k in results}
for (M, N, K, en) in SHAPES:
    key = "%d,%d,%d,%s" % (M, N, K, en)
    row = {k: v.get(key, float("nan")) for k, v in results.items()}
    for k in row:
        tot[k] += row[k]
    best = min((v, k) for k, v in row.items() if k != "auto")[1] if len(row) > 1 else "-"
    print("%-40s %9.1f %9.1f %9.1f %9.1f   %8.0f  %s" % (key, row.get("auto", 0), row.get("bn256", 0), row.get("bn128", 0), row.get("pair", 0),
                                                       2.0 * M * N * K / row.get("auto", 1) / 1e6, best))
print("# totals (us):", {k: round(v, 1) for k, v in tot.items()})

"""Per-kernel times of the fused NHWC UNet ops (csrc/unet_norm.cu) at the shapes of the 512x512 guided-diffusion UNet,
from CUPTI kernel records (torch.profiler), L2 flushed between iterations (512 MB memset), against the measured HBM peak.
Algorithmic bytes: stats pass reads x; apply reads x, writes y; bwd partial reads dy + x; bwd apply reads dy + x, writes dx.
Usage (GPU box):  python tools/bench_unet_kernels.py > profiles/<name>.txt"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

from clip_diffusion_b200 import unet_ops

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = peaks.get("hbm_gbs", 6650.0)
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ITERS = 5


def kernel_times(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(ITERS):
            flush_buf.zero_()
            fn()
        torch.cuda.synchronize()
    out = {}
    for ev in prof.events():
        if ev.device_type.name != "CUDA" or "Memset" in ev.name or "FillFunctor" in ev.name:
            continue
        name = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        r = out.setdefault(name, [0, 0.0])
        r[0] += 1
        r[1] += ev.device_time
    return {k: v[1] / v[0] for k, v in out.items()}


def report(title, times, bytes_of):
    for name, us in times.items():
        nb = next((b for key, b in bytes_of.items() if key in name), None)
        s = "%-44s %-46s %8.1f us" % (title, name[:46], us)
        if nb:
            gbs = nb / us / 1e3
            s += "  %8.1f GB/s  %5.1f%% of %.0f" % (gbs, 100 * gbs / HBM, HBM)
        print(s, flush=True)


print("# mean of %d, CUPTI kernel durations, L2 flushed between iterations; peak = measured HBM copy bandwidth" % ITERS)
shapes = [(1, 128, 512, 512), (1, 256, 512, 512), (1, 256, 256, 256), (1, 512, 256, 256), (1, 256, 128, 128), (1, 512, 64, 64), (1, 1024, 32, 32),
          (1, 2048, 16, 16), (1, 1024, 8, 8)]
for shp in shapes:
    N, C, H, W = shp
    x = torch.randn(shp, device="cuda").half().contiguous(memory_format=torch.channels_last).requires_grad_()
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    ss = torch.randn(N, 2 * C, device="cuda") * 0.1
    dy = torch.randn(shp, device="cuda").half().contiguous(memory_format=torch.channels_last)
    nb = x.numel() * 2
    y = unet_ops.group_norm_nhwc(x, gamma, beta, 32, 1e-5, scale_shift=ss, silu=True)
    t = kernel_times(lambda: unet_ops.group_norm_nhwc(x, gamma, beta, 32, 1e-5, scale_shift=ss, silu=True))
    report("GN+ss+SiLU fwd %s" % (shp,), t, {"gn_stats_partial": nb, "gn_apply_fwd": 2 * nb})
    t = kernel_times(lambda: torch.autograd.grad(y, x, dy, retain_graph=True))
    report("GN+ss+SiLU bwd %s" % (shp,), t, {"gn_bwd_partial": 2 * nb, "gn_apply_bwd": 3 * nb})
for shp in [(1, 128, 512, 512), (1, 256, 256, 256), (1, 512, 64, 64)]:
    N, C, H, W = shp
    a = torch.randn(shp, device="cuda").half().contiguous(memory_format=torch.channels_last)
    b = torch.randn(shp, device="cuda").half().contiguous(memory_format=torch.channels_last)
    bias = torch.randn(C, device="cuda")
    nb = a.numel() * 2
    report("bias_residual_add %s" % (shp,), kernel_times(lambda: unet_ops.bias_residual_add(a, b, bias)), {"bias_residual_add": 3 * nb})
    report("avg_pool2x %s" % (shp,), kernel_times(lambda: unet_ops.avg_pool2x(a)), {"resample2x": nb + nb // 4})
    report("upsample_nearest2x %s" % (shp,), kernel_times(lambda: unet_ops.upsample_nearest2x(a)), {"resample2x": 5 * nb})
    if hasattr(unet_ops, "concat_channels"):
        report("concat_channels %s" % (shp,), kernel_times(lambda: unet_ops.concat_channels(a, b)), {"concat2": 4 * nb})
        c = unet_ops.concat_channels(a, b)
        report("split (concat backward) %s" % (shp,), kernel_times(lambda: unet_ops._split_channels(c, C, C)), {"split2": 4 * nb})

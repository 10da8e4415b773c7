"""Kernel-time table of ONE guidance step (torch.profiler / CUPTI, no serialisation): which kernels the step spends its
device time in.  Usage: python tools/profile_step.py [workload] > profiles/<name>.txt"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from torch.profiler import ProfilerActivity, profile

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
sys.argv = ["bench.py", "--workload", wl, "--steps", "1", "--warmup", "3", "--no-cpu-baseline"]
# run the normal benchmark with the timed region wrapped by the CUDA profiler range, collecting CUPTI kernel records
prof = profile(activities=[ProfilerActivity.CUDA], record_shapes=False)
orig_start, orig_stop = torch.cuda.profiler.start, torch.cuda.profiler.stop
torch.cuda.profiler.start = lambda: prof.__enter__()
torch.cuda.profiler.stop = lambda: prof.__exit__(None, None, None)
os.environ["CG_BENCH_PROFILER_RANGE"] = "1"
bench.main()
OURS = ("gemm_bf16_tn_kernel", "attn_fwd_kernel", "attn_bwd_", "attn_delta_kernel", "ln_fwd_kernel", "ln_bwd_kernel", "tables_kernel",
        "resample_fwd_kernel", "resample_bwd_kernel", "augment_fwd_kernel", "jitter_fwd_kernel", "jitter_bwd_kernel", "affine_bwd_kernel",
        "tv_kernel", "range_kernel", "sph_", "set_cls_kernel", "proj_fwd_kernel", "proj_bwd_kernel", "tokens_to_bf16_kernel", "patchify_kernel",
        "sumsq_nan_kernel", "finalize_kernel", "nanflag_kernel", "gn_stats_partial_kernel", "gn_finalize_", "gn_apply_", "gn_bwd_partial_kernel",
        "bias_residual_add_kernel", "resample2x_kernel", "concat2_kernel", "gemm_bf16_tn_pair_kernel", "hist_kernel", "count_kernel", "::apply_kernel",
        "attn_tc_kernel", "attn_fwd_tc_kernel", "image_losses_kernel", "ms_pool_kernel", "ms_maps_kernel", "ms_scalars_kernel", "ms_grad_kernel", "ms_up_kernel",
        "producer_stats_kernel", "jitter_bwd_sum_kernel", "randn_like_torch_kernel", "sph_loss_sum_kernel")
rows = {}
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    r = rows.setdefault(ev.name, [0, 0.0])
    r[0] += 1
    r[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
total = sum(v[1] for v in rows.values())
print("# workload %s: device time of one timed step = %.3f ms over %d kernel/memcpy records" % (wl, total / 1e3, sum(v[0] for v in rows.values())))
print("# %-7s %-6s %-10s name" % ("share", "calls", "total_us"))
mine = 0.0
for name, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:(10000 if os.environ.get("PROFILE_ALL") else 45)]:
    print("%6.2f%% %6d %10.1f  %s" % (100 * us / total, n, us, name[:150]))
for name, (n, us) in rows.items():
    if any(k in name for k in OURS):
        mine += us
print("# clipguide_b200 kernels: %.3f ms = %.2f%% of device time" % (mine / 1e3, 100 * mine / total))
for name, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    if any(k in name for k in OURS):
        print("#   %6d %10.1f us  %s" % (n, us, name[:110]))

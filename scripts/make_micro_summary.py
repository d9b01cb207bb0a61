"""profiles/<round>_microbench.{json,md} from gpurun_out/<round>_micro.json (bench.py --mode micro)."""
import json, os, shutil, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", f"{R}_micro.json")
d = json.load(open(src))
shutil.copy(src, os.path.join(ROOT, "profiles", f"{R}_microbench.json"))
with open(os.path.join(ROOT, "profiles", f"{R}_microbench.md"), "w") as f:
    f.write(f"# {R} -- kNN + Conv_surface / Conv_layer microbenchmark sweep (BASELINE.json configs[3])\n\n"
            "`python bench.py --mode micro` on one B200: S = 7, C = 128, D = 128 for the feature-space kNN; B clouds of N points with "
            "B*N ~ 32 x 1028; median of CUDA-event times, 256 MB L2 flush before each iteration (not under a profiler).\n"
            f"Roofs: FP32 {d['fp32_peak_tflops']:.1f} TFLOP/s (148 SM x 128 lanes x 2 x 1.965 GHz, derived), HBM {d['hbm_peak_gbs']:.0f} GB/s ({d['peak_source']}).\n"
            "Fractions are ALGORITHMIC work (SURVEY 8d) / time / roof: kNN `N^2(2D+3)+N^2` flops, surface conv `N k S C 8 + N S C`, "
            "layer conv `N k S C 9 + N S C` flops and its compulsory bytes.  The feature-space kNN computes its inner products on the "
            "tensor cores (k <= 63 since round 2), so its fraction of the FP32 (CUDA-core) roof can exceed 100 %.\n\n"
            "| N | B | k | kNN xyz ms | Gpairs/s | % FP32 | kNN feat ms | Gpairs/s | % FP32 | path | surface conv ms | % FP32 | layer conv ms | % FP32 | % HBM | table |\n"
            "|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|---:|---:|---:|---:|---|\n")
    for r in d["rows"]:
        x, ft, s, l = r["knn_xyz"], r["knn_feat_D128"], r["surface_conv"], r["layer_conv"]
        f.write(f"| {r['N']} | {r['B']} | {r['k']} | {x['ms']:.3f} | {x['gpairs_s']:.0f} | {100 * x['fp32_frac']:.1f} | "
                f"{ft['ms']:.3f} | {ft['gpairs_s']:.0f} | {100 * ft['fp32_frac']:.1f} | {'tcgen05' if ft['path'].startswith('tcgen05') else 'fp32'} | "
                f"{s['ms']:.3f} | {100 * s['fp32_frac']:.1f} | {l['ms']:.3f} | {100 * l['fp32_frac']:.1f} | {100 * l['hbm_frac']:.1f} | {l['table']} |\n")
print("ok")

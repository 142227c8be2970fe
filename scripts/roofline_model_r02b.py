"""Per-kernel-class distance to the roofline for one B = 1024, L = 20 LXMERT step of the round-2b build (self-attention inside the
Q|K|V projection, single-word fp16 state), from the committed ncu launch list:

    python scripts/roofline_model_r02b.py [profiles/r02b_launch_summary.json]

Algorithmic FLOPs (2 M N K; attention 4 Sq Sk 64 per head) and algorithmic HBM bytes (every operand read once, every result
written once, weights once per launch) per class, roofline minimum = max(FLOPs / tensor peak, bytes / HBM peak) with the MEASURED
peaks of MEASURED_PEAKS.json, against the measured time (ncu `gpu__time_duration.sum` summed over the class's launches —
cold-cache, serialised, clocks as the power cap leaves them).  No GPU needed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, L, V, H, I, HEADS, F = 1024, 20, 36, 768, 3072, 12, 2048
ML, MV = B * L, B * V


def main():
    summ = json.load(open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02b_launch_summary.json")))
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        tf, hbm = pk["bf16_tflops_sustained"] * 1e12, pk["hbm_gbs"] * 1e9
    except Exception:
        tf, hbm = 1410.2e12, 6454e9
    us = {k["kernel"].split("::")[-1]: (k["us"], k["launches"], k["dram_bytes"]) for k in summ["kernels"]}
    st = lambda rows: rows * 48                      # row statistics [rows, 6, 2] f32
    w = lambda n, k, launches: 2 * n * k * launches   # 16-bit weight read once per launch
    self_l, self_v = 9 + 5, 5 + 5                     # self-attention blocks per stream (language layers + X layers, relational + X)
    rows_out = (self_l + 5) * ML + (self_v + 5) * MV  # attention-output projections: self + cross blocks
    rows_ffn = self_l * ML + self_v * MV
    rows_x = 5 * (ML + MV)                            # cross-attention Q|K|V rows
    att = lambda n, s: 4 * 64 * HEADS * B * n * s * s
    classes = [
        ("Q‖K‖V + self-attention, visual (one kernel)", ["qkv_attn_pair_kernel<3, 5>"],
         2 * self_v * MV * 3 * H * H + att(self_v, V), self_v * MV * (2 * H * 2) + st(self_v * MV) + w(3 * H, H, self_v)),
        ("Q‖K‖V + self-attention, language (one kernel)", ["qkv_attn_pair_kernel<2, 3>"],
         2 * self_l * ML * 3 * H * H + att(self_l, L), self_l * ML * (2 * H * 2) + st(self_l * ML) + w(3 * H, H, self_l)),
        ("cross-attention Q‖K‖V consumer", ["gemm_bf16_pair_kernel<0, __nv_bfloat16, 1, 8, 3>"],
         2 * rows_x * 3 * H * H, rows_x * (H * 2 + 3 * H * 2) + st(rows_x) + w(3 * H, H, 10)),
        ("cross attention (TMA kernel)", ["attention_tma_kernel<2>", "attention_tma_kernel<3>"],
         4 * 64 * HEADS * B * 5 * 2 * L * V, rows_x * (3 * H * 2 + H * 2)),
        ("attention-output producer (K = 768)", ["gemm_bf16_pair_kernel<7, __nv_bfloat16, 1, 8, 3>"],
         2 * rows_out * H * H, rows_out * (H * 2 + 2 * H * 2) + 2 * st(rows_out) + w(H, H, 34)),
        ("FFN1 consumer (erf-GELU)", ["gemm_bf16_pair_kernel<1, __nv_bfloat16, 1, 8, 3>"],
         2 * rows_ffn * I * H, rows_ffn * (H * 2 + I * 2) + st(rows_ffn) + w(I, H, 24)),
        ("FFN2 producer (K = 3072)", ["gemm_bf16_pair_kernel<7, __nv_bfloat16, 1, 8, 2>"],
         2 * rows_ffn * I * H, rows_ffn * (I * 2 + 2 * H * 2) + 2 * st(rows_ffn) + w(H, I, 24)),
        ("VisualFeatEncoder (one kernel, fp32 features)", ["visn_embed_kernel<1>"],
         2 * MV * H * F, MV * (F * 4 + H * 2) + st(MV) + H * F * 4),
    ]
    print(f"peaks: {tf / 1e12:.0f} TFLOP/s (sustained bf16), {hbm / 1e9:.0f} GB/s; one step = {summ['step_kernel_time_us']:.0f} us of kernel "
          f"time, {summ['launches']} launches, {summ['dram_bytes_per_step'] / 1e9:.1f} GB of DRAM traffic\n")
    print("| kernel class | launches | GFLOP | algorithmic GB | ncu DRAM GB | bound | roofline min µs | measured µs | fraction of roofline |")
    print("|---|---|---|---|---|---|---|---|---|")
    tot_min = tot_meas = 0.0
    out = []
    for name, ks, f, b in classes:
        n = sum(us[k][1] for k in ks)
        meas = sum(us[k][0] for k in ks)
        dram = sum(us[k][2] for k in ks)
        t_f, t_b = f / tf * 1e6, b / hbm * 1e6
        tmin = max(t_f, t_b)
        tot_min += tmin
        tot_meas += meas
        bound = f"tensor ({t_f:.0f} / {t_b:.0f} µs)" if t_f >= t_b else f"HBM ({t_f:.0f} / {t_b:.0f} µs)"
        print(f"| {name} | {n} | {f / 1e9:.0f} | {b / 1e9:.2f} | {dram / 1e9:.2f} | {bound} | {tmin:.0f} | {meas:.0f} | {tmin / meas:.2f} |")
        out.append({"class": name, "launches": n, "gflop": f / 1e9, "algorithmic_gb": b / 1e9, "ncu_dram_gb": dram / 1e9,
                    "roofline_min_us": tmin, "measured_us": meas, "frac": tmin / meas})
    print(f"| the eight classes together | | | | | | {tot_min:.0f} | {tot_meas:.0f} | {tot_min / tot_meas:.2f} |")
    if "--json" in sys.argv:
        json.dump({"classes": out, "total_min_us": tot_min, "total_measured_us": tot_meas}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()

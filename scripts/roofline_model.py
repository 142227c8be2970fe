"""Per-kernel-class distance to the roofline for one B = 1024, L = 20 LXMERT step, from the committed ncu launch list:

    python scripts/roofline_model.py [profiles/r01_launch_summary.json]

For each hot kernel class: algorithmic FLOPs (2 M N K; attention 4 Sq Sk 64 per head) and algorithmic HBM bytes (every operand
read once, every result written once, bf16 unless noted; weights counted once per launch), the roofline-minimum time
max(FLOPs / tensor peak, bytes / HBM peak) with the MEASURED peaks of MEASURED_PEAKS.json, and the measured time (ncu
`gpu__time_duration.sum`, summed over the class's launches — cold-cache, serialised, clocks as the power cap leaves them).
No GPU needed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, L, V, H, I, HEADS = 1024, 20, 36, 768, 3072, 12
ML, MV = B * L, B * V


def main():
    summ = json.load(open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01_launch_summary.json")))
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        tf, hbm = pk["bf16_tflops_sustained"] * 1e12, pk["hbm_gbs"] * 1e9
    except Exception:
        tf, hbm = 1410.2e12, 6454e9
    us = {k["kernel"].split("::")[-1]: (k["us"], k["launches"], k["dram_bytes"]) for k in summ["kernels"]}
    # token rows through each linear of the encoder: 9 language layers, 5 relational layers, 5 cross-modality layers
    # (cross-attention + self-attention + FFN on both streams)
    att_l, att_v = 9 + 5 + 5, 5 + 5 + 5            # attention blocks (Q|K|V + output projection) per stream
    ffn_l, ffn_v = 9 + 5, 5 + 5
    rows_att, rows_ffn = att_l * ML + att_v * MV, ffn_l * ML + ffn_v * MV
    n_att, n_ffn = att_l + att_v, ffn_l + ffn_v
    w = lambda n, k, launches: 2 * n * k * launches
    stats = lambda rows: rows * 48
    classes = [
        ("Q‖K‖V consumer", "gemm_bf16_pair_kernel<0, __nv_bfloat16, 1, 8, 3>", 2 * rows_att * 3 * H * H,
         rows_att * H * 2 + stats(rows_att) + rows_att * 3 * H * 2 + w(3 * H, H, n_att)),
        ("attention-output producer", "gemm_bf16_pair_kernel<4, __nv_bfloat16, 1, 8, 3>", 2 * rows_att * H * H,
         rows_att * H * 2 + 2 * rows_att * H * 4 + 2 * stats(rows_att) + w(H, H, n_att)),
        ("FFN1 consumer (erf-GELU)", "gemm_bf16_pair_kernel<1, __nv_bfloat16, 1, 8, 3>", 2 * rows_ffn * I * H,
         rows_ffn * H * 2 + stats(rows_ffn) + rows_ffn * I * 2 + w(I, H, n_ffn)),
        ("FFN2 producer", "gemm_bf16_pair_kernel<4, __nv_bfloat16, 1, 8, 2>", 2 * rows_ffn * I * H,
         rows_ffn * I * 2 + 2 * rows_ffn * H * 4 + 2 * stats(rows_ffn) + w(H, I, n_ffn)),
    ]
    # attention: language self (9 + 5), visual self (5 + 5), cross both ways (5 each); reads Q|K|V rows, writes the context rows
    fl_att = 4 * 64 * HEADS * B * ((9 + 5) * L * L + (5 + 5) * V * V + 5 * 2 * L * V)
    by_att = rows_att * 3 * H * 2 + rows_att * H * 2
    t_att = us["attention_tma_kernel<2>"][0] + us["attention_tma_kernel<3>"][0]
    print(f"peaks: {tf / 1e12:.0f} TFLOP/s (sustained bf16), {hbm / 1e9:.0f} GB/s; one step = {summ['step_kernel_time_us']:.0f} us of kernel time\n")
    print("| kernel class | launches | GFLOP | algorithmic GB | ncu DRAM GB | bound | roofline min µs | measured µs | fraction of roofline |")
    print("|---|---|---|---|---|---|---|---|---|")
    tot_min = tot_meas = 0.0
    rows = [(n, us[k][1], f, b, us[k][2], us[k][0]) for n, k, f, b in classes]
    rows.append(("attention (both kernels)", us["attention_tma_kernel<2>"][1] + us["attention_tma_kernel<3>"][1], fl_att, by_att,
                 us["attention_tma_kernel<2>"][2] + us["attention_tma_kernel<3>"][2], t_att))
    for name, n, f, b, dram, meas in rows:
        t_f, t_b = f / tf * 1e6, b / hbm * 1e6
        tmin = max(t_f, t_b)
        tot_min += tmin
        tot_meas += meas
        print(f"| {name} | {n} | {f / 1e9:.0f} | {b / 1e9:.2f} | {dram / 1e9:.2f} | {'tensor' if t_f >= t_b else 'HBM'} "
              f"({t_f:.0f} / {t_b:.0f} µs) | {tmin:.0f} | {meas:.0f} | {tmin / meas:.2f} |")
    print(f"| the five classes together | | | | | | {tot_min:.0f} | {tot_meas:.0f} | {tot_min / tot_meas:.2f} |")


if __name__ == "__main__":
    main()

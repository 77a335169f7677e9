"""Summarise an `ncu --set full` report of the GEMM launches of one training step (run here, no GPU needed):

    python scripts/ncu_summary.py gpurun_out/prof_gemm.ncu-rep profiles/r01_ncu_full_gemm_<tag>.csv [profiles/r01_traffic.json]

Writes one row per launch with the metrics DESIGN.md / bench.py quote, and (optionally) the DRAM bytes per launch per kernel kind
that bench.py copies into `roofline.traffic`."""
import csv
import io
import json
import subprocess
import sys

KEEP = ["launch__grid_size", "launch__cluster_dim_x", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]


def kind_of(name: str) -> str:
    # template arguments: gemm_tn_kernel<BN, STAGES, EPI, CG, VAR>
    if "gemm_nt_kernel" in name:
        return "gemm_nt_wgrad"
    if "gemm_tn_kernel" in name:
        epi = int(name.split("<")[1].split(">")[0].split(",")[2])
        return {0: "gemm_tn_fwd", 1: "gemm_tn_head", 2: "gemm_tn_dgrad", 6: "gemm_tn_dgrad"}.get(epi, "gemm_tn_other")
    if "tail_kernel" in name:
        return "gemm_tn_head"
    return "other"


def main():
    rep, out_csv = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {}
    for k in KEEP:
        cands = [i for i, h in enumerate(hdr) if h == k or h.endswith("." + k) or h.endswith(k)]
        if cands:
            col[k] = cands[0]
    ki = hdr.index("Kernel Name")
    per_kind = {}
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["ID", "Kernel Name"] + [k for k in KEEP if k in col])
        w.writerow(["", ""] + [units[col[k]] for k in KEEP if k in col])
        for r in body:
            w.writerow([r[0], r[ki]] + [r[col[k]] for k in KEEP if k in col])
            def val(k):
                v = float(r[col[k]].replace(",", ""))
                u = units[col[k]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            kd = per_kind.setdefault(kind_of(r[ki]), {"launches": 0, "bytes": 0.0})
            kd["launches"] += 1
            kd["bytes"] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    print(f"{len(body)} launches -> {out_csv}")
    if len(sys.argv) > 3:
        js = {k: {"launches": v["launches"], "dram_bytes_per_launch": v["bytes"] / v["launches"]} for k, v in per_kind.items()}
        js["source"] = (f"ncu --set full --clock-control none, one training step at B=65536 ({out_csv}): dram__bytes_read.sum + "
                        "dram__bytes_write.sum averaged over the launches of each kernel kind")
        json.dump(js, open(sys.argv[3], "w"), indent=1)
        print(json.dumps(js, indent=1))


if __name__ == "__main__":
    main()

"""Developer tool: turn an `ncu --set full` report of one eager fused step (tools/prof_step.py) into the two artefacts the repo
commits per round: a per-kernel summary table (profiles/rN_ncu_full_step.csv) and the DRAM bytes per launch that bench.py
reports as `roofline.traffic` (profiles/rN_ncu_traffic.json).  Runs on the CPU box:  ncu -i <rep> --page raw --csv | this script.

    ncu -i gpurun_out/prof_step.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv profiles/r2_ncu_full_step.csv profiles/r2_ncu_traffic.json "<note>"
"""
import csv
import json
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
# C-ABI call name (bench.py's ALGO keys) <- kernel name fragment
CALLS = {"ncn_grid_bwd": "grid_bwd_merge_kernel", "ncn_grid_fwd": "grid_fwd_coherent_kernel", "ncn_field_mlp_fwd": "field_mlp_fwd_tc05_kernel",
         "ncn_mlp_bwd_src_fused": "mlp_bwd_tc05_kernel", "ncn_adam_step_groups": "adam_kernel", "ncn_adam_step": "adam_kernel",
         "ncn_grad_sumsq": "sumsq_kernel", "ncn_composite_train_fw_photometric": "composite_train_fw_sw_kernel", "ncn_composite_train_fw_photometric_gt": "composite_train_fw_sw_kernel",
         "ncn_composite_train_fw": "composite_train_fw_sw_kernel", "ncn_composite_train_bw": "composite_train_bw_sw_kernel",
         "ncn_march_train_expand": "march_train_expand_kernel", "ncn_cluster_chain": "kmeans_kernel"}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
    raw, out_csv, out_json, note = sys.argv[1], sys.argv[2], sys.argv[3], (sys.argv[4] if len(sys.argv) > 4 else "")
    rows = list(csv.reader(open(raw)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hi], rows[hi + 1]
    name_i = hdr.index("Kernel Name")
    col_i = {c: hdr.index(c) for c in COLS if c in hdr}
    data = [r for r in rows[hi + 2:] if len(r) == len(hdr)]
    # the LAST occurrence of every kernel = the warmest eager step
    last = {}
    for r in data:
        last[r[name_i]] = r
    order = sorted(last.values(), key=lambda r: -float(r[col_i["gpu__time_duration.sum"]].replace(",", "")) *
                   {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[col_i["gpu__time_duration.sum"]], 1))
    with open(out_csv, "w") as f:
        f.write(f'"# {note}"\n')
        f.write("kernel," + ",".join(f"{c} [{units[i]}]" for c, i in col_i.items()) + ",launches_in_capture\n")
        for r in order:
            n = sum(1 for q in data if q[name_i] == r[name_i])
            f.write('"' + r[name_i][:110].replace('"', "'") + '",' + ",".join(r[i].replace(",", "") for i in col_i.values()) + f",{n}\n")
    traffic = {"_source": f"{out_csv} ({note}); dram__bytes_read.sum + dram__bytes_write.sum per launch"}
    for call, frag in CALLS.items():
        hits = [r for r in data if frag in r[name_i]]
        if not hits:
            continue
        # average over the launches of the last step (calls that launch twice per step: two different template instances)
        names = sorted({r[name_i] for r in hits})
        per = []
        for nm in names:
            r = last[nm]
            per.append(to_bytes(r[col_i["dram__bytes_read.sum"]], units[col_i["dram__bytes_read.sum"]]) +
                       to_bytes(r[col_i["dram__bytes_write.sum"]], units[col_i["dram__bytes_write.sum"]]))
        traffic[call] = {"kernel": " | ".join(n[:60] for n in names), "bytes_per_launch": sum(per) / len(per)}
    json.dump(traffic, open(out_json, "w"), indent=1)
    print(open(out_csv).read())


if __name__ == "__main__":
    main()

"""Reads the ncu reports / bench lines a GPU visit left in gpurun_out/ and writes the tracked summaries under profiles/.
Usage: python tools/summarize_profiles.py <tag>      (e.g. r01_tc)"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                d[k] = (r[hdr.index(k)] + " " + units[hdr.index(k)]).strip()
        res.append(d)
    return res


def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not hi:
        return {}
    hdr = rows[hi[0]]
    data = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))] if len(r) >= len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: 0 for h in names}
    ts = 0
    ops = {}
    for r in data:
        ts += int(r[ix["# Samples"]] or 0)
        for h in names:
            tot[h] += int(r[ix[h]] or 0)
        src = r[ix["Source"]].split()
        op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]] or 0)
    ti = sum(ops.values()) or 1
    return {"samples": ts, "stall_pct": {h: round(100.0 * v / max(ts, 1), 1) for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]},
            "opcode_pct_of_instructions": {k: round(100.0 * v / ti, 1) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]}}


def main():
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    summary = {}
    traffic = {}
    tp = os.path.join(PROF, "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    for name, workload in (("prof_wavenet", "a1_standard"), ("prof_lstm", "lstm_1x16"), ("prof_a2", "a2_full"), ("prof_lstm_tc", None), ("prof_lstm_tc2", None), ("prof_nano", None)):
        rep = os.path.join(OUT, name + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        launches = raw(rep)
        if not launches:
            continue
        summary[name] = {"launches": launches, "sampling": stalls(rep)}
        if workload is None:
            continue
        try:
            rd = float(launches[0]["dram__bytes_read.sum"].split()[0]) * (1e6 if "Mbyte" in launches[0]["dram__bytes_read.sum"] else 1e9 if "Gbyte" in launches[0]["dram__bytes_read.sum"] else 1e3 if "Kbyte" in launches[0]["dram__bytes_read.sum"] else 1)
            wr = float(launches[0]["dram__bytes_write.sum"].split()[0]) * (1e6 if "Mbyte" in launches[0]["dram__bytes_write.sum"] else 1e9 if "Gbyte" in launches[0]["dram__bytes_write.sum"] else 1e3 if "Kbyte" in launches[0]["dram__bytes_write.sum"] else 1)
            traffic[workload] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": launches[0]["kernel"][:60], "from": tag}
            if "wavenet_h_kernel<" in launches[0]["kernel"]:
                traffic[workload]["kernel_choice"] = "tcgen05_fp16_pairs"
            if workload == "a2_full":
                # the bench's A2 step is 256 frames = two 128-frame launches of the tensor-core kernel
                traffic[workload]["launches_per_step"] = 2
                traffic[workload]["dram_bytes_per_step"] = 2 * (rd + wr)
        except Exception as e:
            print("traffic parse failed", e)
    json.dump(summary, open(os.path.join(PROF, tag + "_ncu_summary.json"), "w"), indent=1)
    json.dump(traffic, open(tp, "w"), indent=1)
    for f in os.listdir(OUT):
        if f.startswith("bench_") and f.endswith(".json") and os.path.getsize(os.path.join(OUT, f)) > 0:
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, tag + "_" + f))
    if os.path.exists(os.path.join(OUT, "launches.csv")):
        shutil.copy(os.path.join(OUT, "launches.csv"), os.path.join(PROF, tag + "_launches_a1_standard.csv"))
    for f, dst in (("pytest_gpu.log", "_pytest_gpu.log"), ("parity.txt", "_parity.txt"), ("latency.txt", "_single_stream_latency.txt"),
                   ("model_test.txt", "_model_test.txt"), ("shape_table.txt", "_shape_table.txt"), ("lstm_tc_check.txt", "_lstm_tc_check.txt"),
                   ("host_paths_ab.txt", "_host_paths_ab.txt")):
        if os.path.exists(os.path.join(OUT, f)):
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, tag + dst))
    print(json.dumps({k: v["launches"][0] for k, v in summary.items()}, indent=1)[:3000])


if __name__ == "__main__":
    main()

"""Dev tool (container side): compact, committable summaries of ncu captures brought back in gpurun_out/.

  python tools/summarise_ncu.py rep gpurun_out/r02_lv.ncu-rep profiles/r02_lv_mh_kernel_ncu_full.csv [chains_per_gpu]
      raw page of an `ncu --set full` report -> `metric,unit,launch0,launch1,...` rows, filtered to launch /
      dram / pipe / issue / stall / occupancy metrics (the file bench.py parses for roofline.traffic)
  python tools/summarise_ncu.py generic gpurun_out/r02_generic_metrics.csv profiles/r02_generic_instr.json
      per-launch FP64 instruction counts of the one-chain-per-thread kernels (tools/prof_generic.py) ->
      the JSON bench.py uses for the C2 / C3 FP64-issue fractions
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(r"dram__bytes|gpu__time_duration|launch__|sm__pipe_fp64|sm__inst_executed_pipe_fp64|sm__inst_executed_pipe_(alu|fma|xu|lsu|uniform)"
                  r"|smsp__issue_active|sm__warps_active|smsp__average_warps?_issue_stalled|smsp__warp_issue_stalled|sm__pipe_tensor"
                  r"|smsp__inst_executed\.sum|sm__throughput|lts__t_sector_hit_rate|l1tex__data_bank_conflicts|smsp__pcsamp_warps_issue_stalled"
                  r"|sm__inst_executed_pipe_fp64|smsp__inst_executed_pipe_fp64|shared_(ld|st)_bank_conflict|l1tex__data_pipe_lsu_wavefronts_mem_shared")


def rep(path, out, chains=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, launches = rows[0], rows[1], rows[2:]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
        if chains:
            w.writerow(["chains_per_gpu", ""] + [chains] * len(launches))
        for col, name in enumerate(head):
            if name in ("Kernel Name", "Block Size", "Grid Size") or KEEP.search(name):
                short = name.split(".", 2)[-1] if re.match(r"^[A-Z_]+\.Triage", name) else name
                w.writerow([short, units[col]] + [r[col] for r in launches])
    print(f"{out}: {len(launches)} launches")


def generic(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head = rows[0]
    idx = {k: head.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Value")}
    per = {}
    for r in rows[1:]:
        per.setdefault((r[idx["ID"]], r[idx["Kernel Name"]], r[idx["Grid Size"]]), {})[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
    # tools/prof_generic.py: three launches of 2,000 transitions each; C3 = <2,2,1,*> on 16,384 chains, C2 = <2,2,0,*> on 4,096
    res = {}
    for (i, kern, grid), m in sorted(per.items(), key=lambda kv: int(kv[0][0])):
        name = "C3" if "<2, 2, 1" in kern else "C2"
        chains = 16384 if name == "C3" else 4096
        res[name] = {"kernel": kern, "grid": grid, "chain_steps_per_launch": 2000 * chains,       # last (warm) launch wins
                     "fp64_warp_inst_per_launch": m["sm__inst_executed_pipe_fp64.sum"],
                     "warp_inst_per_launch": m["smsp__inst_executed.sum"],
                     "pipe_fp64_cycles_active_pct": m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"],
                     "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"],
                     "warps_active_pct": m["sm__warps_active.avg.pct_of_peak_sustained_active"],
                     "gpu_time_ms": m["gpu__time_duration.sum"] * 1e-6,
                     "dram_bytes": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]}
    res["source"] = "ncu --metrics ... -k regex:generic_mh_kernel python tools/prof_generic.py (tools/ncu_round2.sh generic)"
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"rep": rep, "generic": generic}[sys.argv[1]](*sys.argv[2:])

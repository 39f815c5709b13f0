#!/usr/bin/env python
"""Summarise one `ncu --set full --import-source on` capture (run here, on the .ncu-rep brought back in gpurun_out/).

  python tools/ncu_summary.py gpurun_out/X.ncu-rep [--json profiles/rNN_bench_ncu.json --views 8 --command "..."]

Prints the SASS-region breakdown used in profiles/*.txt (runs of instructions with equal execution counts: exec/inst,
average active lanes, share of issued warp instructions, share of stall samples, dominant opcodes) and selected raw
metrics; with --json also writes the per-launch figures bench.py reads for `roofline.traffic`.
"""
import argparse, csv, json, subprocess, sys

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--json")
ap.add_argument("--views", type=int, default=8)
ap.add_argument("--command", default="")
ap.add_argument("--head", default="", help="commit (and a word on the kernel) the capture was taken at; bench.py quotes it")
ap.add_argument("--algorithmic-bytes-per-view", type=int, default=543162368)
a = ap.parse_args()

out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
ie, src, at, ss = h.index("Instructions Executed"), h.index("Source"), h.index("Avg. Threads Executed"), h.index("# Samples")
data = [(int(r[ie]), float(r[at]) if r[at] else 0, r[src].strip(), int(r[ss])) for r in rows[2:] if len(r) > ie]
tot, tots = sum(d[0] for d in data), sum(d[3] for d in data)
# fetches really issued: predicated-off lanes (e.g. mixed-label samples that reuse the group's fetch) do not count
pon = h.index("Predicated-On Thread Instructions Executed")
tex_lane_fetches = sum(int(r[pon]) for r in rows[2:] if len(r) > pon and r[src].strip() and
                       (r[src].split()[1] if r[src].strip().startswith("@") else r[src].split()[0]).startswith("TEX"))
print("texture fetches (lanes) %.4e" % tex_lane_fetches)
print("total warp-inst %.3e" % tot, "n sass", len(data), "samples", tots)
i = 0
while i < len(data):
    j = i
    while j + 1 < len(data) and abs(data[j + 1][0] - data[i][0]) <= 0.02 * max(data[i][0], 1) + 1:
        j += 1
    cnt, smp = sum(d[0] for d in data[i:j + 1]), sum(d[3] for d in data[i:j + 1])
    if cnt / tot > 0.004 or smp / tots > 0.01:
        ops = {}
        for d in data[i:j + 1]:
            op = d[2].split()[0] if not d[2].startswith("@") else d[2].split()[1]
            op = op.split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda x: -x[1])[:9]
        print(f"sass[{i:4d}-{j:4d}] n={j - i + 1:3d} exec/inst={data[i][0]:.3e} thr={data[i][1]:4.1f} inst%={cnt / tot * 100:5.1f} "
              f"stall-samples%={smp / tots * 100:5.1f} {top}")
    i = j + 1

raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[-1]
m = {hh: vals[k] for k, hh in enumerate(hdr)}
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active", "l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__f_tex2sm_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]
for w in want:
    if w in m:
        print(f"{w:80s} {m[w]}")


def f(k, scale=1.0):
    try:
        return float(m[k].replace(",", "")) * scale
    except Exception:
        return None


if a.json:
    units = {hh: rows[1][k] for k, hh in enumerate(hdr)} if len(rows) > 2 else {}

    def in_bytes(k):
        v = f(k)
        if v is None:
            return None
        u = units.get(k, "byte").lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)

    def in_ms(k):
        v = f(k)
        u = units.get(k, "ns").lower()
        return None if v is None else v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}.get(u, 1e-6)
    rd, wr = in_bytes("dram__bytes_read.sum"), in_bytes("dram__bytes_write.sum")
    inst, cyc = f("smsp__inst_executed.sum"), f("sm__cycles_elapsed.avg")
    js = {"command": a.command, "head": a.head, "kernel": m.get("Kernel Name", ""), "views_per_launch": a.views,
          "march_dram_bytes_per_launch": (rd or 0) + (wr or 0), "dram_bytes_read": rd, "dram_bytes_write": wr,
          "algorithmic_bytes_per_launch": a.algorithmic_bytes_per_view * a.views, "duration_ms_under_ncu": in_ms("gpu__time_duration.sum"),
          "warp_instructions": inst, "sm_cycles": cyc, "tex_lane_fetches_per_launch": tex_lane_fetches,
          "issue_slot_utilisation": f("sm__issue_active.avg.pct_of_peak_sustained_elapsed", 0.01),
          "shared_pipe_wavefronts_pct": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
          "avg_active_lanes": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
          "l1tex_hit_pct": f("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": f("lts__t_sector_hit_rate.pct"),
          "tex_request_cycles_pct": f("l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed"),
          "pipe_fma_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
          "pipe_alu_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
          "pipe_xu_pct": f("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
          "registers_per_thread": f("launch__registers_per_thread"), "grid": f("launch__grid_size"), "block": f("launch__block_size"),
          "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
          "dram_throughput_pct": f("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed") or f("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
          # what binds: the texture data pipe of the L1TEX unit (wavefronts per cycle), its filter stage, the LSU data pipe (shared memory)
          "l1tex_throughput_pct": f("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
          "l1tex_tex_data_pipe_wavefronts_pct": f("l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed"),
          "l1tex_filter_wavefronts_pct": f("l1tex__f_wavefronts.avg.pct_of_peak_sustained_elapsed"),
          "l1tex_lsu_data_pipe_wavefronts_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
          "tex_output_wavefronts": f("l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum"),
          "tex_requests": f("l1tex__t_requests_pipe_tex_mem_texture.sum"),
          "l2_bytes_per_launch": (f("lts__t_sectors.sum") or 0) * 32.0,
          "l2_throughput_pct": f("lts__throughput.avg.pct_of_peak_sustained_elapsed")}
    if js["duration_ms_under_ncu"]:
        js["l2_GBps"] = js["l2_bytes_per_launch"] / js["duration_ms_under_ncu"] / 1e6
        js["dram_GBps"] = js["march_dram_bytes_per_launch"] / js["duration_ms_under_ncu"] / 1e6
    json.dump(js, open(a.json, "w"), indent=1)
    print("wrote", a.json)

#!/usr/bin/env python
"""tools/ncu_tables.py tag — markdown tables + profiles/traffic.json from gpurun_out/<tag>_prof_{C2,C1M,C4,denoise}.ncu-rep."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
M = {"time": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1hit": "l1tex__t_sector_hit_rate.pct",
     "l2hit": "lts__t_sector_hit_rate.pct", "l1thr": "l1tex__throughput.avg.pct_of_peak_sustained_active",
     "smthr": "sm__throughput.avg.pct_of_peak_sustained_elapsed", "regs": "launch__registers_per_thread",
     "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio", "inst": "smsp__inst_executed.sum",
     "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "fma": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
     "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "bank": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(M.values())], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        k = {"name": d["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")}
        for key, m in M.items():
            if m in d and d[m] != "":
                v = float(d[m].replace(",", ""))
                if key in ("time", "rd", "wr"):
                    v *= UNIT.get(u[m], 1.0)
                k[key] = v
        res.append(k)
    return res


def table(rows, labels):
    cols = [("time", "gpu time (us)", "{:.1f}"), ("dram", "DRAM rd / wr (MB)", None), ("dram_pct", "DRAM % of peak", "{:.2f}"), ("l1hit", "L1 hit %", "{:.1f}"),
            ("l2hit", "L2 hit %", "{:.1f}"), ("l1thr", "L1TEX thr. %", "{:.1f}"), ("smthr", "SM thr. %", "{:.1f}"), ("regs", "regs", "{:.0f}"),
            ("warps", "warps active %", "{:.1f}"), ("issue", "issue slots busy %", "{:.1f}"), ("lanes", "**lanes / instruction**", "**{:.1f}**"),
            ("inst", "warp instructions (M)", None)]
    out = ["| metric | " + " | ".join(labels) + " |", "|---|" + "---|" * len(rows)]
    for key, title, fmt in cols:
        cells = []
        for r in rows:
            if key == "dram":
                cells.append(f"{r['rd']/1e6:.1f} / {r['wr']/1e6:.1f}")
            elif key == "inst":
                cells.append(f"{r['inst']/1e6:.1f}")
            else:
                cells.append(fmt.format(r[key]) if key in r else "-")
        out.append(f"| {title} | " + " | ".join(cells) + " |")
    return "\n".join(out)


G = os.path.join(ROOT, "gpurun_out")
c2 = load(os.path.join(G, f"{tag}_prof_C2.ncu-rep"))
c1m = load(os.path.join(G, f"{tag}_prof_C1M.ncu-rep"))
c4 = load(os.path.join(G, f"{tag}_prof_C4.ncu-rep"))
dn = load(os.path.join(G, f"{tag}_prof_denoise.ncu-rep"))
print("### C2 (81 922 triangles, 1080p): the seven kernels of one dispatch\n")
print(table(c2, ["`k_primary`", "`k_shade_primary`", "`k_trace_persistent<0>` secondary", "`<1>` shadow depth 0", "`k_shade_secondary`", "`<1>` shadow depth 1", "`k_resolve`"]))
tot = sum(r["time"] for r in c2)
print(f"\nframe (serialised under ncu): {tot:.0f} us; traversal {sum(c2[i]['time'] for i in (0,2,3,5))/tot*100:.1f} %, shading {sum(c2[i]['time'] for i in (1,4,6))/tot*100:.1f} %; "
      f"DRAM of the three shading kernels {sum(c2[i]['rd']+c2[i]['wr'] for i in (1,4,6))/1e6:.0f} MB = {sum(c2[i]['rd']+c2[i]['wr'] for i in (1,4,6))/2073600:.0f} B per pixel and sample\n")
tl = ["`k_primary`", "`k_trace_persistent<0>` secondary", "`<1>` shadow depth 0", "`<1>` shadow depth 1"]
print("### C1M (1 310 722 triangles, traversal section ~440 MB, 1080p)\n")
print(table(c1m, tl))
print("\n### C4 (512 instances of one 20 480-triangle BLAS = 10.5 M triangles, 1080p)\n")
print(table(c4, tl))
print("\n### DenoiseCompositor (1080p, maxKernelSize 12)\n")
print(table(dn, ["`k_denoise<0>` (H)", "`k_denoise<1>` (V + composite)"]))
for r, n in zip(dn, ("H", "V")):
    print(f"\npass {n}: pipes ALU {r['alu']:.0f} % / FMA {r['fma']:.0f} % / LSU {r['lsu']:.0f} %, shared-memory bank conflicts {r['bank']/1e3:.0f} k; "
          f"{r['inst']*32/2073600:.0f} thread instructions per pixel")
traffic = {"source": f"profiles/r2_ncu.md (ncu --set full --clock-control none, one pixel band per dispatch (RT_BANDS=1), dram__bytes_read.sum + dram__bytes_write.sum per launch; 1080p)",
           "dram_bytes_per_launch": {"primary (k_primary)": c2[0]["rd"] + c2[0]["wr"],
                                     "secondary_incoherent (k_trace_persistent<0>)": c2[2]["rd"] + c2[2]["wr"],
                                     "shadow (k_trace_persistent<1>, 2 launches/frame)": (c2[3]["rd"] + c2[3]["wr"] + c2[5]["rd"] + c2[5]["wr"]) / 2,
                                     "C1M secondary_incoherent (k_trace_persistent<0>)": c1m[1]["rd"] + c1m[1]["wr"],
                                     "C4 secondary_incoherent (k_trace_persistent<0>)": c4[1]["rd"] + c4[1]["wr"]},
           "issue_counters": {"primary (k_primary)": {"warp_instructions": c2[0]["inst"], "lanes_per_instruction": c2[0]["lanes"]},
                              "secondary_incoherent (k_trace_persistent<0>)": {"warp_instructions": c2[2]["inst"], "lanes_per_instruction": c2[2]["lanes"]},
                              "shadow (k_trace_persistent<1>, 2 launches/frame)": {"warp_instructions": (c2[3]["inst"] + c2[5]["inst"]) / 2,
                                                                                     "lanes_per_instruction": (c2[3]["inst"] * c2[3]["lanes"] + c2[5]["inst"] * c2[5]["lanes"]) / (c2[3]["inst"] + c2[5]["inst"])},
                              "C1M secondary_incoherent (k_trace_persistent<0>)": {"warp_instructions": c1m[1]["inst"], "lanes_per_instruction": c1m[1]["lanes"]}}}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)

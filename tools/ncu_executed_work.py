#!/usr/bin/env python3
"""Executed work of ONE pairing step from an ncu report: wide-MAC instructions (IMAD.WIDE*) and DRAM bytes.

    python tools/ncu_executed_work.py REPORT.ncu-rep LOG2_BATCH BUILD_ID [OUT.json]

REPORT must hold `ncu --set full --import-source on` captures of the launches of one step WITHOUT the two-stream
split of the final exponentiation (k_pairing<1> + 6 x k_fe_stage; the six k_fe_batch_inv launches are counted when
captured, else their work is added from the arithmetic of fp_batch_inv).  Writes profiles/executed_work.json, which
bench.py reads for `roofline.executed_macs_per_pairing` and `roofline.traffic` -- numbers of the shipped build taken
from counters, not typed into the source.
"""
import collections
import csv
import json
import os
import subprocess
import sys

rep, log2, build = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "executed_work.json")
n = 1 << log2


def page(name):
    """REPORT is an .ncu-rep (read through ncu) or the prefix of exported pages PREFIX.raw.csv / PREFIX.source.csv[.gz]
    (`ncu -i REP --page raw|source --csv`, exported on the GPU box when the report is too large to bring back)."""
    if rep.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    else:
        import gzip
        path = "%s.%s.csv" % (rep, name)
        text = gzip.open(path + ".gz", "rt").read() if os.path.exists(path + ".gz") else open(path).read()
    return list(csv.reader(text.splitlines()))


# ---- raw page: one row per captured launch
raw = page("raw")
hdr = raw[0]
ix = {h: i for i, h in enumerate(hdr)}
units = raw[1]


def bytes_of(row, key):
    v, u = float(row[ix[key]]), units[ix[key]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]


launches = []
for r in raw[2:]:
    if len(r) < len(hdr):
        continue
    launches.append({"kernel": r[ix["Kernel Name"]].split("(")[0], "ms": float(r[ix["gpu__time_duration.sum"]]) * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[units[ix["gpu__time_duration.sum"]]],
                     "dram_bytes": bytes_of(r, "dram__bytes_read.sum") + bytes_of(r, "dram__bytes_write.sum"),
                     "warp_inst": float(r[ix["smsp__inst_executed.sum"]]),
                     "local_loads": float(r[ix["sass__inst_executed_local_loads"]]), "local_stores": float(r[ix["sass__inst_executed_local_stores"]]),
                     "fmaheavy_pct": float(r[ix["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]]),
                     "registers": int(float(r[ix["launch__registers_per_thread"]]))})

# ---- source page: per-instruction executed counts, one block per captured launch (same order)
src = page("source")
blocks, cur = [], None
for r in src:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1] if len(r) > 1 else "", "hdr": None, "wide": 0, "narrow_imad": 0, "all": 0}
        blocks.append(cur)
        continue
    if cur is None:
        # single-kernel reports have no "Kernel Name" rows: the first row is the header
        cur = {"name": "", "hdr": None, "wide": 0, "narrow_imad": 0, "all": 0}
        blocks.append(cur)
    if cur["hdr"] is None:
        if "Source" in r and "Instructions Executed" in r:
            cur["hdr"] = {h: i for i, h in enumerate(r)}
        continue
    h = cur["hdr"]
    if len(r) <= h["Instructions Executed"] or not r[h["Instructions Executed"]].strip().isdigit():
        continue
    ex = int(r[h["Instructions Executed"]])
    s = r[h["Source"]]
    cur["all"] += ex
    if "IMAD.WIDE" in s:
        cur["wide"] += ex
    elif "IMAD" in s:
        cur["narrow_imad"] += ex
blocks = [b for b in blocks if b["all"]]
assert len(blocks) == len(launches), (len(blocks), len(launches))
for l, b in zip(launches, blocks):
    l["wide_mac_warp_inst"] = b["wide"]
    l["narrow_imad_warp_inst"] = b["narrow_imad"]

by = collections.OrderedDict()
for l in launches:
    k = by.setdefault(l["kernel"], {"launches": 0, "ms": 0.0, "dram_bytes": 0.0, "wide_macs": 0.0, "narrow_imads": 0.0, "warp_inst": 0.0,
                                    "local_loads": 0.0, "local_stores": 0.0, "registers": l["registers"], "fmaheavy_pct_time_weighted": 0.0})
    k["launches"] += 1
    k["ms"] += l["ms"]
    k["dram_bytes"] += l["dram_bytes"]
    k["wide_macs"] += 32.0 * l["wide_mac_warp_inst"]
    k["narrow_imads"] += 32.0 * l["narrow_imad_warp_inst"]
    k["warp_inst"] += l["warp_inst"]
    k["local_loads"] += l["local_loads"]
    k["local_stores"] += l["local_stores"]
    k["fmaheavy_pct_time_weighted"] += l["fmaheavy_pct"] * l["ms"]
for k in by.values():
    k["fmaheavy_pct_time_weighted"] /= max(k["ms"], 1e-9)

wide = sum(k["wide_macs"] for k in by.values())
dram = sum(k["dram_bytes"] for k in by.values())
note = []
if not any("batch_inv" in name for name in by):
    # Montgomery's trick over runs of 16: 3 * 15 products + one 380-squaring / ~228-product Fermat ladder per run, six inversions per pairing
    add = 6 * (45 + 608) * 300.0 / 16 * n
    wide += add
    note.append("k_fe_batch_inv not captured: its %.0f wide MACs per pairing added from the arithmetic of fp_batch_inv" % (add / n))
res = {"build": build, "log2_batch": log2, "report": os.path.basename(rep),
       "executed_wide_macs_per_pairing": wide / n, "dram_bytes_per_pairing": dram / n, "dram_bytes_per_step": dram,
       "step_ms_under_ncu": sum(k["ms"] for k in by.values()), "kernels": by, "notes": note}
with open(out, "w") as f:
    json.dump(res, f, indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "kernels"}, indent=1))
for name, k in by.items():
    print("%-40s x%d  %.3f ms  wideMAC/pairing %.0f  dram %.2f GB  local ld/st %.1fM/%.1fM  fmaheavy %.1f%%  regs %d" % (
        name[:40], k["launches"], k["ms"], k["wide_macs"] / n, k["dram_bytes"] / 1e9, k["local_loads"] / 1e6, k["local_stores"] / 1e6,
        k["fmaheavy_pct_time_weighted"], k["registers"]))

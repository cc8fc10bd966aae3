"""Per-tile instruction profile of a scan kernel from an `ncu --page source --csv` export:
    python scripts/ncu_regions.py gpurun_out/ncu_c2_single_source.csv [n_tiles] [listing.txt]
Prints totals, the opcode mix and the contiguous SASS regions by executions per tile."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 37450.0
hdr, data = rows[1], rows[2:]
ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
iw, ie = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")
tot = sum(int(r[ia]) for r in data)
print(rows[0][1][:110])
print(f"warp-instr {tot}  per tile {tot / tiles:.1f}  SASS lines {len(data)}  samples {sum(int(r[ismp]) for r in data)}")
print(f"smem wavefronts {sum(int(r[iw]) for r in data)} ({sum(int(r[iw]) for r in data) / tiles:.1f} per tile), excessive {sum(int(r[ie]) for r in data)}")
op = collections.Counter()
for r in data:
    s = re.sub(r"^@!?U?P\w+\s+", "", r[isrc].strip())
    op[s.split()[0].split(".")[0]] += int(r[ia])
print("opcodes per tile: " + "  ".join(f"{k} {v / tiles:.1f}" for k, v in op.most_common(18)))
base = int(data[0][0], 16)
regs, cur = [], None
lst = open(sys.argv[3], "w") if len(sys.argv) > 3 else None
for r in data:
    ex, idx = int(r[ia]) / tiles, (int(r[0], 16) - base) // 16
    if lst:
        lst.write(f"{idx:5d} {ex:7.2f} {int(r[ismp]):4d} {int(r[iw]) / tiles:6.2f}  {r[isrc].strip()}\n")
    k = round(ex, 1)
    if cur and abs(cur["k"] - k) <= 0.15:
        cur["n"] += 1; cur["s"] += int(r[ismp]); cur["e"] = idx; cur["w"] += ex; cur["wf"] += int(r[iw]) / tiles
    else:
        cur = {"k": k, "n": 1, "s": int(r[ismp]), "b": idx, "e": idx, "w": ex, "wf": int(r[iw]) / tiles}
        regs.append(cur)
for g in regs:
    if g["w"] >= 3:
        print(f"[{g['b']:5d}-{g['e']:5d}] exec/tile {g['k']:5.1f} lines {g['n']:4d} instr/tile {g['w']:7.1f} smem wf/tile {g['wf']:6.1f} samples {g['s']}")

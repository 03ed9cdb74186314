#!/usr/bin/env python3
"""Dynamic opcode histogram of one kernel from `cuobjdump -sass`: every backward branch closes a loop whose trip count is given
on the command line (in order of appearance), and the histogram is weighted accordingly.  Prints the issue-model estimate
(sum count / measured lane-rate, profiles/pipe_bench_r01.txt) next to the counts.
usage: sass_dyn.py file.o kernel_substring trip1,trip2,...  [--div N]   (N = permutations per thread etc.)"""
import re, subprocess, sys, collections
RATE = {"IADD3": 183, "VIADDMNMX": 116, "VIADD": 116, "IMAD": 85, "IMAD.IADD": 85, "IMAD.MOV": 85, "IMAD.SHL": 85, "IMAD.X": 85, "IMAD.HI": 43,
        "IMAD.WIDE": 25, "SHF": 89, "LEA": 89, "LEA.HI": 89, "LOP3": 116, "IADD3.X": 89, "PRMT": 89, "SEL": 116, "ISETP": 116, "MOV": 116}
def main():
    obj, sub, trips = sys.argv[1], sys.argv[2], [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 and sys.argv[3] else []
    div = 1.0
    if "--div" in sys.argv: div = float(sys.argv[sys.argv.index("--div") + 1])
    txt = subprocess.check_output(["cuobjdump", "-sass", obj], text=True)
    cur, ins = None, []
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m: cur = m.group(1); continue
        if cur is None or sub not in cur: continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            addr = int(m.group(1), 16); body = re.sub(r"^@!?U?P\w+\s+", "", m.group(2).strip())
            ins.append((addr, body))
    if not ins: sys.exit("kernel not found")
    w = [1.0] * len(ins); idx = {a: i for i, (a, _) in enumerate(ins)}; k = 0
    for i, (a, b) in enumerate(ins):
        m = re.match(r"BRA(\.U)?\s+.*?(0x[0-9a-f]+)", b)
        if m and int(m.group(2), 16) <= a and int(m.group(2), 16) in idx:
            t = trips[k] if k < len(trips) else 1; k += 1
            for j in range(idx[int(m.group(2), 16)], i + 1): w[j] *= t
            print(f"loop {k}: {m.group(2)}..{a:#x} ({i + 1 - idx[int(m.group(2), 16)]} instr) x{t}", file=sys.stderr)
    h = collections.Counter()
    for (a, b), x in zip(ins, w):
        op = b.split()[0]; parts = op.split(".")
        key = parts[0]
        if key == "IMAD" and len(parts) > 1 and parts[1] in ("HI", "WIDE", "MOV", "IADD", "SHL", "X"): key = "IMAD." + parts[1]
        if key in ("IADD3", "LEA") and len(parts) > 1 and parts[1] in ("X", "HI"): key += "." + parts[1]
        h[key] += x
    tot = sum(h.values()); clk = 0
    for kk, v in h.most_common():
        r = RATE.get(kk); c = v / r if r else 0; clk += c
        print(f"{kk:12s} {v / div:9.1f}  {c / div:7.2f} clk" + ("" if r else "  (unrated)"))
    print(f"total {tot / div:.1f} instr, issue floor {tot / div / 128:.2f} clk, additive pipe model {clk / div:.2f} clk")
main()

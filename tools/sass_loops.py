"""List the loops of one kernel in an object file with their SASS opcode histograms (instruction-count budgeting)."""
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
for part in txt.split("Function :")[1:]:
    name = part.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if pat not in dem.replace(" ", "").replace("(int)", "").replace("(bool)", ""):
        continue
    print(dem[:120])
    lines = [l for l in part.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
    addr = lambda l: int(re.match(r"\s+/\*([0-9a-f]{4,5})\*/", l).group(1), 16)
    A = [addr(l) for l in lines]
    print("  total SASS instructions:", len(lines))
    for i, l in enumerate(lines):
        m = re.search(r"BRA\S*\s+(?:.*?)0x([0-9a-f]+)", l)
        if m and int(m.group(1), 16) < A[i]:
            tgt = int(m.group(1), 16)
            body = [x for x in lines if tgt <= addr(x) <= A[i]]
            if not any("MUFU.EX2" in b for b in body) or len(body) > 1200:
                continue
            ops = {}
            for b in body:
                mm = re.search(r"\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", b)
                if mm:
                    ops[mm.group(2)] = ops.get(mm.group(2), 0) + 1
            print(f"  loop {hex(tgt)}..{hex(A[i])}: {len(body)} instr", sorted(ops.items(), key=lambda x: -x[1]))
    break

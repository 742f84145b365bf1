"""Per-kernel SASS instruction counts of libzelll_b200.so (cuobjdump -sass): total instructions and the
mnemonics that matter for the pair kernels.  `python scripts/sass_summary.py [substring] [--dump DIR]`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "zelll_b200", "libzelll_b200.so")
KEYS = ["FFMA2", "FFMA", "FADD", "FMUL", "DADD", "DMUL", "DFMA", "DSETP", "MUFU", "SHF", "VIMNMX3", "VIMNMX", "VOTE", "POPC",
        "REDUX", "FLO", "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "UBLKCP", "SYNCS", "BAR", "CALL", "BRA", "SEL", "ISETP"]


def main():
    argv = sys.argv[1:]
    dump = None
    if "--dump" in argv:
        k = argv.index("--dump")
        dump = argv[k + 1]
        del argv[k:k + 2]
    want = argv
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    funcs = {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                funcs[name] = body
            name, body = m.group(1), []
        elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            body.append(line)
    if name:
        funcs[name] = body
    for name, body in funcs.items():
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if want and not all(w in dem for w in want):
            continue
        cnt = collections.Counter()
        for line in body:
            m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                op = m.group(1)
                for k in KEYS:
                    if op == k or op.startswith(k):
                        cnt[k] += 1
                        break
        print(f"{len(body):6d}  {dem[:150]}")
        print("        " + " ".join(f"{k}={v}" for k, v in cnt.items() if v))
        if dump:
            os.makedirs(dump, exist_ok=True)
            safe = re.sub(r"[^A-Za-z0-9]+", "_", dem)[:120]
            with open(os.path.join(dump, safe + ".sass"), "w") as f:
                f.write("\n".join(body) + "\n")


main()

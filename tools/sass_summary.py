"""Per-kernel SASS opcode counts of the shipped library (cuobjdump -sass), written to profiles/<tag>_sass_opcodes.txt.
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA),
UTMALDG / UTMASTG = tensor-map TMA, LDGSTS = cp.async, SYNCS = mbarrier ops."""
import re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "REDG", "ATOMS", "MUFU", "LDG", "STG", "LDS", "STS"]

def main(tag):
    so = ROOT / "dualpixelface_b200" / "libdpf_sm100.so"
    txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
    rows = []
    for p in re.split(r"\n\s*Function : ", txt)[1:]:
        name = p.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", "-p", name], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
        rows.append((dem, len(re.findall(r"/\*[0-9a-f]{4}\*/", p)), {o: len(re.findall(r"\b" + o + r"[\.\s]", p)) for o in OPS}))
    tot = {o: sum(r[2][o] for r in rows) for o in OPS}
    out = ROOT / "profiles" / f"{tag}_sass_opcodes.txt"
    with open(out, "w") as f:
        f.write(f"# cuobjdump -sass dualpixelface_b200/libdpf_sm100.so ({so.stat().st_size} bytes, {len(rows)} kernels); opcode sites per kernel\n")
        f.write("# totals: " + ", ".join(f"{o} {tot[o]}" for o in OPS) + "\n")
        f.write(f"{'kernel':84s} {'inst':>6s} " + " ".join(f"{o:>7s}" for o in OPS) + "\n")
        for dem, n, c in sorted(rows, key=lambda r: (-r[2]['UTCHMMA'], r[0])):
            f.write(f"{dem[:84]:84s} {n:6d} " + " ".join(f"{c[o]:7d}" for o in OPS) + "\n")
    print(open(out).read()[:3000])

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")

"""Static evidence for profiles/: per kernel of libnasrec_b200.so the ptxas resource line (registers, spills, shared
memory; from the -Xptxas -v build logs in csrc/obj/*.log) and the counts of the SASS mnemonics that show which hardware
path a kernel uses (tcgen05: UTCHMMA / UTCBAR / LDTM / STTM / UTCATOM*, TMA: UTMALDG / UTMASTG / UBLKCP, mbarrier: SYNCS,
clusters / DSMEM: UCGABAR / MAPA-style LDS.*CLUSTER / MEMBAR).  No GPU needed.

    python tools/sass_summary.py > profiles/rNN_sass.md
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nasrec_b200", "lib", "libnasrec_b200.so")
KEYS = [("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
        ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG|\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("UCGABAR", r"\bUCGABAR"),
        ("LDS/LD cluster", r"\bLD[SG]?\.[A-Z0-9.]*CLUSTER|\bLDS.*\bcluster|\bLD\.E.*\.SHARED"), ("REDG/ATOMG", r"\bREDG|\bATOMG|\bRED\.|\bATOM\."),
        ("FFMA", r"\bFFMA"), ("LDG", r"\bLDG"), ("STG", r"\bSTG")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(n):
    n = n.replace("(anonymous namespace)::", "")
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    n = re.sub(r"cub::CUB_\w+::", "cub::", n)
    return n if len(n) < 110 else n[:107] + "..."


def resources():
    res = {}
    for log in glob.glob(os.path.join(ROOT, "nasrec_b200", "csrc", "obj", "*.log")):
        cur = None
        for line in open(log):
            m = re.search(r"Compiling entry function '([^']+)'", line)
            if m:
                cur = m.group(1)
                res[cur] = {"spill": "0/0"}
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                res[cur]["stack"] = int(m.group(1))
                res[cur]["spill"] = f"{m.group(2)}/{m.group(3)}"
            m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", line)
            if m and cur:
                res[cur]["regs"] = int(m.group(1))
                res[cur]["smem"] = int(m.group(3) or 0)
    return res


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        body = line.split("/*", 2)
        ins = body[1].split("*/", 1)[1] if len(body) > 1 and "*/" in body[1] else line
        if not re.search(r"[A-Z]{3}", ins):
            continue
        counts[cur]["_n"] += 1
        for k, rx in KEYS:
            if re.search(rx, ins):
                counts[cur][k] += 1
    res = resources()
    dm = demangle(list(counts))
    print("# Static SASS / ptxas summary of `libnasrec_b200.so` (sm_100a)\n")
    print("`python tools/sass_summary.py` -- `cuobjdump -sass` mnemonic counts and the `-Xptxas -v` lines of the build that made")
    print("the library in this tree.  Counts are static instructions (loops count once).  Columns: tcgen05 MMA issue")
    print("(`UTC*MMA`), tcgen05 commit / barrier (`UTCBAR`), TMEM loads / stores (`LDTM` / `STTM`), TMA tensor loads")
    print("(`UTMALDG`), mbarrier operations (`SYNCS`), cluster barrier (`UCGABAR`).\n")
    hdr = ["kernel", "SASS instr", "regs", "spill st/ld B", "static smem B", "UTC*MMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "SYNCS", "UCGABAR", "FFMA", "LDG", "STG", "RED/ATOM"]
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    rows = []
    for k, c in counts.items():
        r = res.get(k, {})
        rows.append((short(dm.get(k, k)), c["_n"], r.get("regs", "-"), r.get("spill", "-"), r.get("smem", "-"), c["UTCHMMA"], c["UTCBAR"],
                     c["LDTM"], c["STTM"], c["UTMALDG"], c["SYNCS"], c["UCGABAR"], c["FFMA"], c["LDG"], c["STG"], c["REDG/ATOMG"]))
    rows.sort(key=lambda r: (-(r[5] > 0), -(r[9] > 0), r[0]))
    for r in rows:
        print("| `" + str(r[0]) + "` | " + " | ".join(str(x) for x in r[1:]) + " |")
    spilled = [r[0] for r in rows if r[3] not in ("0/0", "-")]
    print(f"\n{len(rows)} kernels; kernels with register spills: {', '.join('`'+s+'`' for s in spilled) if spilled else 'none'}.")


if __name__ == "__main__":
    sys.exit(main())

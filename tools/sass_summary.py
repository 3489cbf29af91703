"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UTMAREDG (TMA tensor copies), UBLKCP (cp.async.bulk), SYNCS (mbarrier), and
the legacy tensor path HMMA (must be absent).

    python tools/sass_summary.py pytorch-gan_b200/b200gan/libb200gan.so > profiles/rN_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

KEYS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA",
        "ATOMS", "RED", "LDGSTS")


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    kern = None
    counts = collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern)
            counts[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and kern:
            op = m.group(1)
            counts[kern]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[kern][k + ("." + op.split(".", 1)[1] if k.startswith("UTMA") and "." in op else "")] += 1
    print(f"SASS mnemonic summary of {path} (sm_100a); kernels with tcgen05 / TMA / bulk-copy instructions first\n")
    def weight(c):
        return -(sum(v for k, v in c.items() if k.startswith(("UTC", "LDTM", "UTMA", "UBLKCP"))))
    for kern, c in sorted(counts.items(), key=lambda kv: weight(kv[1])):
        tags = ", ".join(f"{k} x{v}" for k, v in sorted(c.items()) if k != "_total")
        print(f"{kern[:110]:110s} {c['_total']:6d} instr  {tags}")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("\nwhole library: " + ", ".join(f"{k} x{v}" for k, v in sorted(tot.items()) if k != "_total"))
    print("HMMA (legacy mma.sync path) instructions:", tot.get("HMMA", 0))


if __name__ == "__main__":
    main(sys.argv[1])

"""Top stall-sample instructions of one kernel from `ncu -i X.ncu-rep --page source --csv` output."""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    end = next((i for i in range(start + 1, len(rows)) if rows[i] and rows[i][0] in ("Address", "Kernel Name")), len(rows))
    body = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
    col = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[col["# Samples"]] or 0) for r in body)
    print("total samples", tot, "instructions", len(body))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ranked = sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]
    for r in ranked:
        n = int(r[col["# Samples"]] or 0)
        st = sorted(((int(r[col[s]] or 0), s) for s in stall_cols), reverse=True)[:2]
        print(f"{100 * n / max(tot, 1):5.1f}%  {r[col['Source']].strip()[:70]:70s}  " + " ".join(f"{s}={v}" for v, s in st if v))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)

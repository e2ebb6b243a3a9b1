"""Print the instructions with the most warp-stall samples from an `ncu --page source --csv` dump."""
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except (TypeError, ValueError):
        return 0


def report(name, hdr, data, topn):
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(num(r[i_s]) for r in data) or 1
    print("=====", name[:90])
    print("total samples", tot, "instructions", len(data))
    agg = {}
    for r in data:
        for h in stalls:
            agg[h] = agg.get(h, 0) + num(r[hdr.index(h)])
    s = sum(agg.values()) or 1
    print("stall mix:", {k: round(100 * v / s, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for r in sorted(data, key=lambda r: -num(r[i_s]))[:topn]:
        st = {h: num(r[hdr.index(h)]) for h in stalls}
        main_st = sorted(((k, v) for k, v in st.items() if v), key=lambda kv: -kv[1])[:3]
        print(f"{num(r[i_s]):6d} {100 * num(r[i_s]) / tot:5.1f}% ex={r[i_ex]:>8} {r[i_src][:72]:72s} {main_st}")


def main(path, topn=30):
    rows = list(csv.reader(open(path)))
    name, hdr, data = "", None, []
    for r in rows:
        if r and r[0] == "Kernel Name":
            if hdr and data:
                report(name, hdr, data, topn)
            name, hdr, data = r[1], None, []
        elif "Source" in r and "# Samples" in r:
            hdr = r
        elif hdr and len(r) == len(hdr):
            data.append(r)
    if hdr and data:
        report(name, hdr, data, topn)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)

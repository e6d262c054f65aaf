"""List the hottest SASS instructions (warp-stall samples) of one kernel of an ncu report.

    python tools/ncu_sass_hot.py gpurun_out/X.ncu-rep <kernel-id> [top]
"""
import csv
import subprocess
import sys


def main():
    rep, kid = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
    if kid != "all":   # "all": single-kernel report (prints the first kernel)
        cmd += ["--kernel-id", ":::%s" % kid]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[0][:2])
    hdr = rows[h]
    idx = {n: i for i, n in enumerate(hdr)}
    data = []
    for r in rows[h + 1:]:
        if r and r[0] == "Kernel Name":
            break  # next kernel of the report
        if len(r) == len(hdr) and r[0] != "Address":
            data.append(r)
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    print("instructions", len(data), "total samples", tot)
    top = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:top_n]
    for i in sorted(top):
        r = data[i]
        st = {n[6:]: int(r[idx[n]]) for n in hdr if n.startswith("stall_") and "Not Issued" not in n and int(r[idx[n]]) > 0}
        print("%5d %-72s %7s %5.1f%% exec=%-8s %s" % (i, r[idx["Source"]][:72], r[idx["# Samples"]],
                                                     100.0 * int(r[idx["# Samples"]]) / max(tot, 1), r[idx["Instructions Executed"]],
                                                     dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])))


if __name__ == "__main__":
    main()

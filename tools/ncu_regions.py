"""Stall samples of an ncu report split into code regions given as name:start:end instruction
indices (see the marker list it prints first).  usage: ncu_regions.py report.ncu-rep [name:a:b ...]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr, data = rows[1], rows[2:]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    keys = ("SETMAXREG", "BAR.SYNC", "EXIT", "LDTM", "ATOMG", "UTCBAR", "SYNCS.ARRIVE")
    print("markers:", [(i, r[isrc].split()[0 if not r[isrc].startswith("@") else 1]) for i, r in enumerate(data)
                       if any(k in r[isrc] for k in keys)])
    tot = sum(int(r[isamp]) for r in data)
    tote = sum(int(r[iex]) for r in data)
    for spec in sys.argv[2:]:
        name, a, b = spec.split(":")
        a, b = int(a), int(b)
        s = sum(int(r[isamp]) for r in data[a:b])
        e = sum(int(r[iex]) for r in data[a:b])
        st = {h: sum(int(r[hdr.index(h)] or 0) for r in data[a:b]) for h in stalls}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:7]
        print("%-10s samples %.3f exec %.3f | " % (name, s / tot, e / tote) +
              ", ".join("%s %.3f" % (k[6:], v / max(s, 1)) for k, v in top))


if __name__ == "__main__":
    main()

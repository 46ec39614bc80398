#!/usr/bin/env python3
"""Per-launch table (kernel, duration, DRAM bytes read / written) from an `ncu --metrics ... --csv` log."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iid, ik, im, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
L = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6, "msecond": 1e3, "usecond": 1, "nsecond": 1e-3, "second": 1e6}.get(u, 1)
    L.setdefault(r[iid], {"k": r[ik].split("(")[0].split("<")[0].split("::")[-1]})[r[im]] = v
for i, d in L.items():
    rd, wr, t = d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0), d.get("gpu__time_duration.sum", 0)
    print(f"{i:>4} {d['k']:28s} {t:10.1f} us  read {rd/1e9:8.3f} GB  write {wr/1e9:8.3f} GB  {(rd+wr)/max(t,1e-9)/1e3:7.1f} GB/s")

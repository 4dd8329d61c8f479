"""ncu launch-list CSV (--metrics gpu__time_duration.sum --csv) -> per-kernel summary CSV (launches, total us, share)."""
import collections, csv, re, sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot = collections.Counter()
cnt = collections.Counter()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void ", "", r[ki])
    name = re.sub(r"\(.*$", "", name).replace("gb::", "")
    name = name[:100]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
with open(dst, "w") as f:
    f.write("# per-launch times are cold-cache and serialised (ncu replays): compare SHARES, not absolutes\n")
    f.write(f"# total {total / 1e3:.2f} ms over {sum(cnt.values())} launches\n")
    f.write("kernel,launches,total_us,share\n")
    for k, v in tot.most_common():
        f.write(f"{k},{cnt[k]},{v:.1f},{v / total:.3f}\n")
print(open(dst).read()[:3000])

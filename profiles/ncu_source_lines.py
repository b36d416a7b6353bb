"""Maps the per-instruction stall samples of an `ncu --set full --import-source on` report to source lines.
usage: ncu_source_lines.py <report.ncu-rep> <library.so built from the profiled tree> > summary.txt
(ncu -i <rep> --page source --csv gives per-SASS-instruction samples; nvdisasm --print-line-info gives the line of every
instruction of the same binary)"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, so = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cub)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
cur = func = None
off2line = {}
for line in dis.splitlines():
    m = re.match(r"^\.text\.(\S+):", line)
    if m:
        func = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.search(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if m and func and "fill_packed_kernelILi16" in func:
        off2line[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode().splitlines()))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
base = int(data[0][0], 16)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(collections.Counter)
tot = collections.Counter()
for r in data:
    li = off2line.get(int(r[0], 16) - base)
    if li is None:
        continue
    key = li[0]
    n = int(r[ci["# Samples"]] or 0); e = int(r[ci["Instructions Executed"]] or 0)
    agg[key]["samples"] += n; agg[key]["inst"] += e; tot["samples"] += n; tot["inst"] += e
    if re.match(r"(LDL|STL)", li[1]):
        tot["local_inst"] += e
    for s_ in stalls:
        v = int(r[ci[s_]] or 0); agg[key][s_] += v; tot[s_] += v
print(f"kernel fill_packed_kernel<16>: {tot['samples']} warp-stall samples, {tot['inst']} warp instructions executed, "
      f"{100.0 * tot['local_inst'] / tot['inst']:.2f} % of them local-memory (spill) loads / stores")
print("stall reasons, % of all samples:", {k[6:]: round(100.0 * v / tot["samples"], 1) for k, v in tot.items() if k.startswith("stall_") and v > 0.005 * tot["samples"]})
print("source lines by samples (file:line, % of samples, % of executed instructions, top stall reasons in % of the line's samples):")
srcs = {}
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:40]:
    top = sorted(((s_, c[s_]) for s_ in stalls if c[s_]), key=lambda x: -x[1])[:3]
    text = ""
    path = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", key[0])
    if os.path.exists(path):
        srcs.setdefault(path, open(path).read().splitlines())
        if key[1] - 1 < len(srcs[path]):
            text = srcs[path][key[1] - 1].strip()[:90]
    print(f"  {key[0]}:{key[1]:<5d} {100.0 * c['samples'] / tot['samples']:5.2f} % {100.0 * c['inst'] / tot['inst']:5.2f} %  "
          f"{[(s_[6:], round(100.0 * v / c['samples'])) for s_, v in top]}  | {text}")

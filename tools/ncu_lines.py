"""Aggregate an ncu SASS-level source page per CUDA source line.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-mangled-substring> [top]

ncu's CSV source page lists SASS instructions in order but without line numbers; nvdisasm -g on the
cubin extracted from the (unchanged) libtda_b200.so lists the same instructions in the same order
with '//## File ... line N' markers.  The two are joined by instruction position."""
import csv, os, re, subprocess, sys, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "tda_eeg_audio_b200", "libtda_b200.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
lines = []  # (line_no) per instruction
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin") or f.count("-") > 0:
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kern not in dis:
        continue
    in_k, cur = False, None
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            in_k = kern in ln
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
    break
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = hdr.index("Instructions Executed")
cs = hdr.index("Warp Stall Sampling (All Samples)")
body = rows[hdr_i + 1:]
if len(body) != len(lines):
    print(f"warning: {len(body)} SASS rows in report vs {len(lines)} in cubin (rebuilt since profiling?)")
agg = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for r, l in zip(body, lines):
    a = agg[l]
    a[0] += int(r[ci]); a[1] += int(r[cs])
    tot_i += int(r[ci]); tot_s += int(r[cs])
src = {}
print(f"total instructions executed {tot_i:,}  stall samples {tot_s:,}")
print(f"{'file:line':28s} {'inst%':>7s} {'stall%':>7s}  source")
for l, (ni, ns) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if l:
        p = os.path.join(ROOT, "tda_eeg_audio_b200", "csrc", l[0])
        if p not in src and os.path.exists(p):
            src[p] = open(p).read().splitlines()
        if p in src and l[1] - 1 < len(src[p]):
            text = src[p][l[1] - 1].strip()[:90]
    print(f"{(l[0] + ':' + str(l[1])) if l else '?':28s} {100 * ni / tot_i:7.2f} {100 * ns / max(tot_s, 1):7.2f}  {text}")

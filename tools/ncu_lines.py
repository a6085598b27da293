"""Per-source-line instruction / stall-sample breakdown of one kernel from an .ncu-rep (SASS page joined with
nvdisasm line info of the library's cubin).  usage: ncu_lines.py report.ncu-rep kernel_substring [min_pct]
BPE_LIB = the library the report was taken with (default: the in-tree build; the join needs the very same SASS)."""
import collections, csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, ksub = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("BPE_LIB") or os.path.join(ROOT, "bpe_tokenizer_b200", "libbpe_b200.so")], cwd=tmp, capture_output=True)
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, "bpe_b200.sm_100a.cubin")], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and ksub in l)
cur, insts = None, []
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append((m.group(2), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = rows[1]; body = [r for r in rows[2:] if len(r) == len(hdr)]
ii, ti, wi = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
assert len(body) == len(insts), (len(body), len(insts))
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = samp = 0
for r, (txt, cur) in zip(body, insts):
    e, t, s = int(r[ii]), int(r[ti]), int(r[wi])
    a = agg[cur]; a[0] += e; a[1] += t; a[2] += s; tot += e; samp += s
src = {}
print("total warp instructions", tot, "samples", samp)
for cur, v in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if v[0] >= tot * minp / 100 or v[2] >= samp * minp / 100:
        f, ln = cur if cur else ("?", 0)
        if f not in src:
            try: src[f] = open(f).read().split("\n")
            except Exception: src[f] = []
        text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
        print("%-16s:%4d inst %5.1f%% thr %4.1f samp %5.1f%% | %s" % (os.path.basename(f)[:16], ln, v[0] * 100 / tot, v[1] / max(1, v[0]), v[2] * 100 / samp, text))

"""Per-source-line view of an ncu report (--import-source on): warp instructions executed, stall samples and
average active threads per CUDA source line, per captured kernel launch.
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top] [launch-index ...]"""
import collections, csv, os, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = set(int(x) for x in sys.argv[3:])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.OrderedDict()
launch, cur_file, seen_files = -1, None, set()
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1])
        if cur_file in seen_files or launch < 0:   # a file repeats -> next captured launch
            launch += 1
            seen_files = set()
        seen_files.add(cur_file)
        continue
    if r[0] == "Function Name":
        fn = r[1]
        continue
    if r[0].isdigit() and len(r) > 10:
        num = lambda x: int(x) if x.isdigit() else 0
        a = agg.setdefault((launch, fn[:46], cur_file, int(r[0]), r[1].strip()[:100]), [0, 0, 0])
        a[0] += num(r[6]); a[1] += num(r[7]); a[2] += num(r[8])
for k in sorted(set(key[0] for key in agg)):
    if only and k not in only:
        continue
    items = [(v[1], key, v) for key, v in agg.items() if key[0] == k]
    tot = sum(i[0] for i in items) or 1
    tots = sum(i[2][0] for i in items) or 1
    tott = sum(i[2][2] for i in items)
    print(f"=== launch {k}: {items[0][1][1]}  warp instr {tot}  thread instr {tott}  avg threads {tott / tot:.2f}")
    for ex, key, v in sorted(items, reverse=True)[:top]:
        print(f"{100 * ex / tot:5.1f}% instr {100 * v[0] / tots:5.1f}% stall  thr {v[2] / max(1, ex):4.1f}  {key[2]}:{key[3]}  {key[4]}")

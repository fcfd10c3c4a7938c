"""Instructions executed / stall samples per CUDA source line of one kernel of an ncu report (needs -lineinfo + --import-source on):
python scratch/ncu_lines.py report.ncu-rep <kernel-index> [top]"""
import csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-id', ':::' + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur_file = ''; agg = {}
for r in rows:
    if r and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r and r[0] == 'Function Name': print(r[1][:120]); continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == '-':      # a source line (aggregated over its SASS)
        ie = float(r[hdr.index('Instructions Executed')]); sm = float(r[hdr.index('# Samples')])
        k = (cur_file, int(r[0]))
        a = agg.setdefault(k, [0.0, 0.0, r[1]]); a[0] += ie; a[1] += sm
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print('total warp-instructions %.0f, samples %.0f' % (tot, tots))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%-14s %4d  inst %5.1f%%  samples %5.1f%% | %s' % (k[0][:14], k[1], 100 * a[0] / tot, 100 * a[1] / max(tots, 1), a[2].strip()[:110]))

"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the committed summaries under profiles/.

  python profiles/make_summary.py <stage.ncu-rep> <aux.ncu-rep> <launches.csv> <chunk_sites> [tag]

Writes profiles/<tag>_stage_tc_final_ncu.txt (per-launch metrics of the stage kernels + the other kernels of the step +
the launch list by kernel) and profiles/<tag>_traffic.json (dram__bytes_read+write of the stage kernels, per site).
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max']
ROLE = {('0', '0'): 'RB4 per-site rows (exits on the device: chunk is dense)', ('0', '1'): 'RB4 stage-1 lattice',
        ('0', '2'): 'RB4 edge pseudo-sites (rows gathered from the stem tables)', ('1', '0'): 'C_RB4 per-site (exits on the device)',
        ('1', '1'): 'C_RB4 stage 2, pre-pooled single-row loader'}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    stage, aux, launches, chunk = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    tag = sys.argv[5] if len(sys.argv) > 5 else "r01"
    with open(os.path.join(ROOT, "profiles", tag + "_stage_tc_final_ncu.txt"), "w") as f:
        f.write("ncu --set full --clock-control none --import-source on -k regex:k_stage_tc ... python bench.py --steps 1 --warmup 1 "
                "--sites-per-step %d\n" % chunk)
        f.write("(one %d-site dense chunk; k_stage_tc<MODE,FM>: MODE 0=RB4 1=C_RB4, FM 0=plain 1=lattice 2=edge gather)\n\n" % chunk)
        hdr, units, rows = raw(stage)
        tot, per = 0.0, []
        for r in rows:
            d = dict(zip(hdr, r))
            nm = d['Kernel Name']
            m = re.search(r'<(\d), (\d)>', nm)
            f.write("== %s   [%s]\n" % (nm, ROLE.get(m.groups(), '') if m else ''))
            for k in KEYS:
                if k in d:
                    f.write("  %-84s %s %s\n" % (k, d[k], units[hdr.index(k)]))
            f.write("\n")
            b = (float(d['dram__bytes_read.sum']) + float(d['dram__bytes_write.sum'])) * 1e6
            tot += b
            per.append({"kernel": nm, "us": float(d['gpu__time_duration.sum']), "dram_bytes": b})
        json.dump({"source": "ncu --set full --clock-control none, profiles/%s_stage_tc_final_ncu.txt (one %d-site dense chunk, %d k_stage_tc "
                             "launches incl. the per-site variants that exit on the device)" % (tag, chunk, len(per)),
                   "chunk_sites": chunk, "launches": len(per), "dram_bytes_total": tot, "dram_bytes_per_launch": tot / len(per),
                   "dram_bytes_per_site": tot / chunk,
                   "note": "bench.py scales per-site DRAM bytes of the stage kernels to its own launch count", "per_launch": per},
                  open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w"), indent=1)
        f.write("---- other kernels of the step (same command, -k regex:k_tail|k_local_mlp_tc|k_dense_tables|k_edge_pool)\n\n")
        hdr, units, rows = raw(aux)
        for r in rows:
            d = dict(zip(hdr, r))
            f.write("== %s\n" % d['Kernel Name'])
            for k in KEYS:
                if k in d:
                    f.write("  %-84s %s %s\n" % (k, d[k], units[hdr.index(k)]))
            f.write("\n")
        rows = list(csv.reader(open(launches)))
        hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
        h, data = rows[hi], rows[hi + 1:]
        ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
        agg = collections.OrderedDict()
        for r in data:
            if len(r) <= vi:
                continue
            try:
                v = float(r[vi].replace(',', ''))
            except ValueError:
                continue
            if r[ui] == 'ns':
                v /= 1000
            a = agg.setdefault(r[ki][:64], [0, 0.0])
            a[0] += 1
            a[1] += v
        t = sum(a[1] for a in agg.values())
        f.write("---- launch list (ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2): profiles/%s_launches_final.csv\n" % tag)
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('  %-66s n=%4d  %9.1f us  %5.1f%%\n' % (k, a[0], a[1], 100 * a[1] / t))
        share = 100 * sum(a[1] for k, a in agg.items() if 'k_stage_tc' in k) / t
        f.write("  stage kernels (k_stage_tc, all roles): %.1f%% of the step under ncu\n" % share)
    print("stage share %.1f%%, DRAM bytes per site %.0f" % (share, tot / chunk))


if __name__ == "__main__":
    main()

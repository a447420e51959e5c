#!/usr/bin/env python
"""Turns `ncu -i X.ncu-rep --page raw --csv` + `--page source --csv` of tools/profile_conv.py Dv.dc2 (three launches: fprop,
dgrad, wgrad) into profiles/r02_ncu_dv_dc2.txt and profiles/r02_traffic.json (DRAM bytes per launch + the csrc sha that
bench.py compares before it reports `roofline.traffic`).  usage: ncu_summary.py raw.csv source.csv"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

LAYER = sys.argv[3] if len(sys.argv) > 3 else "Dv.dc2"
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'l1tex__m_l1tex2xbar_write_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active']
names = {0: 'tc_conv_fprop:' + LAYER, 1: 'tc_conv_dgrad:' + LAYER, 2: 'tc_conv_wgrad:' + LAYER}


def tobytes(v, u):
    return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)


out, traffic = [], {}
for k, r in enumerate(rows[2:]):
    out.append('== launch %d: %s   (%s)' % (k, r[hdr.index('Kernel Name')], names.get(k)))
    d = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            out.append('   %-70s %s %s' % (w, r[i], units[i]))
            d[w] = (r[i], units[i])
    tot = tobytes(*d['dram__bytes_read.sum']) + tobytes(*d['dram__bytes_write.sum'])
    traffic[names[k]] = tot
    out.append('   %-70s %.1f MB' % ('DRAM bytes per launch (read + write)', tot / 1e6))
out += ['', 'Role attribution (--page source, warp stall samples): how many of the single-thread producer / MMA warps\' samples',
        'sit on an mbarrier try_wait (waiting for a free stage / for data) against issuing instructions.']
seen = set()
for sec in open(sys.argv[2]).read().split('"Kernel Name",')[1:]:
    rr = list(csv.reader(('"Kernel Name",' + sec).splitlines()))
    kn = rr[0][1][:90]
    if kn in seen:
        continue
    seen.add(kn)
    h = rr[1]
    isrc, ismp = h.index('Source'), h.index('# Samples')
    data = [(int(r[ismp] or 0), r[isrc].strip()) for r in rr[2:] if len(r) > ismp]
    imma = [j for j, dd in enumerate(data) if 'UTCHMMA' in dd[1]]
    itma = [j for j, dd in enumerate(data) if 'UTMALDG' in dd[1]]

    def region(a, b):
        t = sum(dd[0] for dd in data[a:b])
        w = sum(dd[0] for j, dd in enumerate(data[a:b]) if 'TRYWAIT' in dd[1] or ('BRA' in dd[1] and j > 0 and 'TRYWAIT' in data[a + j - 1][1]))
        return t, w
    pt, pw = region(max(0, itma[0] - 150), itma[-1] + 40)
    mt, mw = region(max(0, imma[0] - 100), imma[-1] + 40)
    out.append('  %s' % kn)
    out.append('     total samples %d; producer code %d samples, %d on `empty` try_wait; MMA-issuer code %d samples, %d on `full` try_wait'
               % (sum(dd[0] for dd in data), pt, pw, mt, mw))
head = ['ncu --set full --clock-control none --import-source on -k regex:tc_ -s 6 -c 3   on   python tools/profile_conv.py ' + LAYER,
        '(third round of fprop / dgrad / wgrad of BASELINE config 2\'s %s; csrc sha %s; one B200; ncu serialises the' % (LAYER, bench.csrc_sha()),
        ' launches and the SMs run ~1.7 GHz under it, so durations are longer than bench.py\'s: read shares and byte counts)', '']
tag = LAYER.lower().replace('.', '_')
open(os.path.join(ROOT, 'profiles', 'r02_ncu_%s.txt' % tag), 'w').write('\n'.join(head + out) + '\n')
tpath = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
old = {}
if os.path.exists(tpath):
    with open(tpath) as f:
        old = json.load(f)
if old.get('csrc_sha') != bench.csrc_sha():
    old = {}
old.update(traffic)
old['csrc_sha'] = bench.csrc_sha()
old['source'] = 'profiles/r02_ncu_<layer>.txt (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch)'
json.dump(old, open(tpath, 'w'), indent=1)
print('\n'.join(out[-8:]))

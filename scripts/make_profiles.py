"""Builds the tracked summaries under profiles/ from the raw artefacts a gpurun session left in gpurun_out/
(scripts/collect_profiles.sh).  Run here (no GPU needed): python scripts/make_profiles.py"""
import csv, io, json, os, subprocess, collections, re
R = 'r1'
G = 'gpurun_out'
P = 'profiles'
os.makedirs(P, exist_ok=True)

# 1. bench lines
for name in ('bench_%s.json' % R, 'bench_ref_%s.json' % R):
    src = os.path.join(G, name)
    if os.path.exists(src):
        line = open(src).read().strip().splitlines()[-1]
        open(os.path.join(P, name), 'w').write(json.dumps(json.loads(line), indent=1) + '\n')

# 2. ncu launch list of the bench command -> per-kernel share
rows = []
with open(os.path.join(G, 'launches_%s.csv' % R)) as f:
    txt = f.read()
txt = txt[txt.index('"ID"'):]
for r in csv.DictReader(io.StringIO(txt)):
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        val = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        us = val / 1000.0 if unit in ('ns', 'nsecond') else (val if unit in ('us', 'usecond') else val * 1000.0)
        rows.append((r['Kernel Name'], us))
with open(os.path.join(P, 'launches_%s.csv' % R), 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-soak\n')
    f.write('# per-launch device time, cold cache and serialised by the profiler: compare SHARES, not absolutes\n')
    f.write('launch,kernel,duration_us\n')
    for i, (k, us) in enumerate(rows):
        f.write('%d,"%s",%.3f\n' % (i, k, us))
agg = collections.OrderedDict()
for k, us in rows:
    short = re.sub(r'\(.*', '', k).split('::')[-1]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += us
det = {k: v for k, v in agg.items() if k.startswith('det_') and 'compact' not in k and 'gather' not in k}
tot = sum(v[1] for v in det.values())

# 3. full-set captures: key metrics per kernel
raw = subprocess.run(['ncu', '-i', os.path.join(G, 'prof_all_%s.ncu-rep' % R), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic']
seen = {}
traffic = {}
def to_bytes(v, u):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
with open(os.path.join(P, 'ncu_full_%s.txt' % R), 'w') as f:
    f.write('# ncu --set full --clock-control none --import-source on  (scripts/collect_profiles.sh); first captured launch of each kernel\n')
    for r in rr[2:]:
        name = re.sub(r'\(.*', '', r[hdr.index('Kernel Name')]).split('::')[-1]
        full = r[hdr.index('Kernel Name')]
        if name in seen:
            continue
        seen[name] = 1
        f.write('\n=== %s\n' % full[:120])
        for w in want:
            if w in hdr:
                f.write('  %-70s %s %s\n' % (w, r[hdr.index(w)], units[hdr.index(w)]))
        rd = to_bytes(r[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_read.sum')])
        wr = to_bytes(r[hdr.index('dram__bytes_write.sum')], units[hdr.index('dram__bytes_write.sum')])
        traffic[name] = int(rd + wr)
tj = {'det_stream_kernel': traffic.get('det_stream_reg_kernel<20, 128>', traffic.get('det_stream_reg_kernel')),
      'by_kernel': traffic}
for k, v in traffic.items():
    if k.startswith('det_stream'):
        tj['det_stream_kernel'] = v
    if k.startswith('target_stream'):
        tj['target_stream_kernel'] = v
json.dump(tj, open(os.path.join(P, 'traffic.json'), 'w'), indent=1)

# 4. hottest source lines per kernel
with open(os.path.join(P, 'ncu_source_%s.txt' % R), 'w') as f:
    for k in ('det_stream', 'det_sort', 'det_nms', 'target_stream', 'target_match'):
        out = subprocess.run(['python', 'scripts/ncu_source.py', os.path.join(G, 'prof_all_%s.ncu-rep' % R), k, '14'], capture_output=True, text=True).stdout
        f.write(out + '\n')

# 5. SASS evidence
sass = subprocess.run(['cuobjdump', '-sass', 'dspnet_b200/libdspmb.so'], capture_output=True, text=True).stdout
cur = None
counts = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = re.sub(r'^_ZN5dspmb\d+_GLOBAL__N__[0-9a-f_]+\w*?\d\d', '', m.group(1))
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)', line)
    if m and cur:
        op = m.group(1)
        for key in ('LDG.E.NA.128', 'LDG.E.NA.64', 'STG.E.NA.128', 'LDG.E.128', 'LDG.E.64', 'STG.E.128', 'UBLKCP', 'SYNCS', 'LDS.128', 'MUFU.EX2', 'DFMA', 'DMUL', 'DADD', 'MATCH', 'VOTE', 'SHFL', 'ATOMS', 'ATOMG', 'RED', 'BAR.SYNC', 'HMMA', 'UTCHMMA'):
            if op.startswith(key):
                counts[cur][key] += 1
with open(os.path.join(P, 'sass_%s.txt' % R), 'w') as f:
    f.write('# cuobjdump -sass dspnet_b200/libdspmb.so (sm_100a): opcode counts per kernel for the mnemonics that matter on this path\n')
    f.write('# (no tensor-core opcodes anywhere: nothing here is a dense contraction; UBLKCP = cp.async.bulk TMA copy)\n')
    for k, c in counts.items():
        short = re.search(r'(prior_kernel|det_[a-z_]+kernel|target_[a-z_]+kernel|nms_[a-z_]+kernel|test_[a-z_]+kernel)(I[\w]*?E)?', k)
        short = (short.group(1) + (short.group(2) or '')) if short else k
        f.write('%-28s %s\n' % (short, ' '.join('%s=%d' % kv for kv in sorted(c.items()))))
print('launch-list shares (detection step):', {k: '%.1f%%' % (100 * v[1] / tot) for k, v in det.items()})
print('traffic', tj)

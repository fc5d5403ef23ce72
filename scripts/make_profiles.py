"""Builds the tracked summaries under profiles/ from the raw artefacts a gpurun session left in gpurun_out/
(scripts/collect_profiles.sh).  Run here (no GPU needed): python scripts/make_profiles.py [round tag, default r2]"""
import collections, csv, io, json, os, re, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else 'r2'
G, P = 'gpurun_out', 'profiles'
os.makedirs(P, exist_ok=True)

# 1. bench lines (one per workload + the reference arms)
for name in sorted(os.listdir(G)):
    if re.match(r'bench_.*_%s\.json$' % R, name) or name == 'bench_ref_%s.json' % R:
        txt = open(os.path.join(G, name)).read().strip()
        if txt:
            open(os.path.join(P, name), 'w').write(json.dumps(json.loads(txt.splitlines()[-1]), indent=1) + '\n')

# 2. ncu launch lists of the bench command -> per-kernel share of the step
def launch_list(src, dst, cmd):
    txt = open(src).read()
    txt = txt[txt.index('"ID"'):]
    rows = []
    for r in csv.DictReader(io.StringIO(txt)):
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            val, unit = float(r['Metric Value'].replace(',', '')), r['Metric Unit']
            us = val / 1000.0 if unit in ('ns', 'nsecond') else (val if unit in ('us', 'usecond') else val * 1000.0)
            rows.append((r['Kernel Name'], us))
    agg = collections.OrderedDict()
    for k, us in rows:
        short = re.sub(r'\(.*', '', k).split('::')[-1].replace('void ', '')
        short = re.sub(r'<.*', '', short)
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for k, v in agg.items() if k.startswith(('det_', 'target_')) and 'gather' not in k and 'compact' not in k)
    with open(dst, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv %s\n' % cmd)
        f.write('# per-launch device time, cold cache and serialised by the profiler: compare SHARES, not absolutes\n')
        f.write('# share of the operator kernels: ' + ', '.join('%s %.1f%% (%d launches, %.1f us avg)' % (k, 100 * v[1] / tot, v[0], v[1] / v[0])
                                                                for k, v in agg.items() if k.startswith(('det_', 'target_'))) + '\n')
        f.write('launch,kernel,duration_us\n')
        for i, (k, us) in enumerate(rows):
            f.write('%d,"%s",%.3f\n' % (i, k, us))
    return agg
shares = {}
if os.path.exists(os.path.join(G, 'launches_%s.csv' % R)):
    shares['detection'] = launch_list(os.path.join(G, 'launches_%s.csv' % R), os.path.join(P, 'launches_%s.csv' % R),
                                      'python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-soak --no-e2e')
if os.path.exists(os.path.join(G, 'launches_target_%s.csv' % R)):
    shares['target'] = launch_list(os.path.join(G, 'launches_target_%s.csv' % R), os.path.join(P, 'launches_target_%s.csv' % R),
                                   'python bench.py --workload target --steps 5 --warmup 3 --no-cpu-baseline --no-soak --no-e2e')

# 3. full-set captures: key metrics per kernel + DRAM traffic per launch
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__waves_per_multiprocessor',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic']
def to_bytes(v, u):
    return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
traffic = {}
with open(os.path.join(P, 'ncu_full_%s.txt' % R), 'w') as f:
    f.write('# ncu --set full --clock-control none --import-source on (scripts/collect_profiles.sh); first captured launch of each kernel\n')
    for rep in ('prof_all_%s' % R, 'prof_heads_%s' % R, 'prof_nms_%s' % R):
        path = os.path.join(G, rep + '.ncu-rep')
        if not os.path.exists(path):
            continue
        raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rr = list(csv.reader(io.StringIO(raw)))
        hdr, units = rr[0], rr[1]
        seen = set()
        for r in rr[2:]:
            full = r[hdr.index('Kernel Name')]
            name = re.sub(r'\(.*', '', full).split('::')[-1].replace('void ', '')
            if name in seen:
                continue
            seen.add(name)
            f.write('\n=== %s   [%s]\n' % (full[:140], rep))
            for w in want:
                if w in hdr:
                    f.write('  %-70s %s %s\n' % (w, r[hdr.index(w)], units[hdr.index(w)]))
            for i, h in enumerate(hdr):
                if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
                    try:
                        v = float(r[i])
                    except ValueError:
                        continue
                    if v > 8:
                        f.write('  STALL %-64s %.1f %%\n' % (h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))
            rd = to_bytes(r[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_read.sum')])
            wr = to_bytes(r[hdr.index('dram__bytes_write.sum')], units[hdr.index('dram__bytes_write.sum')])
            traffic[re.sub(r'<.*', '', name)] = int(rd + wr)
tj = {'by_kernel': traffic, 'source': 'ncu --set full capture of round %s (dram__bytes_read.sum + dram__bytes_write.sum, one launch)' % R}
for k, v in traffic.items():
    if k.startswith('det_stream_bulk'):
        tj['det_stream_kernel'] = v
    if k.startswith('det_stream_heads'):
        tj['det_stream_heads_kernel'] = v
    if k.startswith(('target_stream', 'target_match', 'det_sort', 'det_pair', 'nms_cull', 'nms_resolve')):
        tj[k] = v
json.dump(tj, open(os.path.join(P, 'traffic.json'), 'w'), indent=1)

# 4. hottest source lines per kernel
with open(os.path.join(P, 'ncu_source_%s.txt' % R), 'w') as f:
    for rep, kerns in (('prof_all_%s' % R, ('det_stream', 'det_sort', 'det_pair', 'target_stream', 'target_match')),
                       ('prof_heads_%s' % R, ('det_stream_heads',)), ('prof_nms_%s' % R, ('nms_cull', 'nms_resolve'))):
        path = os.path.join(G, rep + '.ncu-rep')
        if not os.path.exists(path):
            continue
        for k in kerns:
            out = subprocess.run([sys.executable, 'scripts/ncu_source.py', path, k, '14'], capture_output=True, text=True).stdout
            f.write(out + '\n')
            if k in ('det_pair', 'target_match'):
                out = subprocess.run([sys.executable, 'scripts/ncu_phases.py', path, k], capture_output=True, text=True).stdout
                f.write('-- %s by phase (between its %%globaltimer stamps, code-layout order)\n' % k + out + '\n')

# 5. SASS: a real listing of the roofline kernels + opcode counts for every kernel
sass = subprocess.run(['cuobjdump', '-sass', 'dspnet_b200/libdspmb.so'], capture_output=True, text=True).stdout
blocks, cur, name = collections.OrderedDict(), None, None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = name.replace('dspmb::(anonymous namespace)::', '')
        cur = blocks.setdefault(name, [])
        continue
    if cur is not None:
        cur.append(line)
listing = [k for k in blocks if re.match(r'void det_stream_bulk_kernel<20, 128, 2, true, true, true>|void target_stream_kernel<2, 21, true>|void det_stream_heads_kernel<21, true>', k)]
with open(os.path.join(P, 'sass_%s.txt' % R), 'w') as f:
    f.write('# cuobjdump -sass dspnet_b200/libdspmb.so (sm_100a only).  Part 1: opcode counts per kernel for the mnemonics that matter on this\n'
            '# path (no tensor-core opcodes anywhere: nothing here is a dense contraction; UTMALDG / UBLKCP = cp.async.bulk(.tensor) TMA copies, UTMAPF = TMA L2 prefetch, SYNCS = mbarrier,\n'
            '# UCGABAR / ACQBULK = cluster barrier).  Part 2: the full listing of the three streaming kernels.\n\n')
    keys = ('LDG.E.NA.128', 'LDG.E.NA.64', 'LDG.E.NA', 'STG.E.NA.128', 'LDG.E.128', 'LDG.E.64', 'STG.E.128', 'UBLKCP', 'UTMALDG', 'UTMAPF', 'UBLKPF', 'SYNCS', 'UCGABAR', 'LDS.128',
            'MUFU.EX2', 'DFMA', 'DMUL', 'DADD', 'MATCH', 'VOTE', 'SHFL', 'REDUX', 'ATOMS', 'ATOMG', 'RED', 'BAR.SYNC', 'HMMA', 'UTCHMMA')
    for k, lines in blocks.items():
        c = collections.Counter()
        for line in lines:
            m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)', line)
            if m:
                for key in keys:
                    if m.group(1).startswith(key):
                        c[key] += 1
                        break
        f.write('%-70s %s\n' % (k[:70], ' '.join('%s=%d' % kv for kv in sorted(c.items()))))
    for k in listing:
        f.write('\n\n===================== %s =====================\n' % k)
        for line in blocks[k]:
            m = re.search(r'^\s+/\*([0-9a-f]+)\*/\s+(.*?)\s*;?\s*/\*', line)
            if m:
                f.write('/*%s*/ %s\n' % (m.group(1), m.group(2).rstrip(' ;')))
print({k: {kk: round(vv[1] / vv[0], 1) for kk, vv in v.items() if kk.startswith(('det_', 'target_'))} for k, v in shares.items()})
print('traffic', {k: v for k, v in tj.items() if k not in ('by_kernel', 'source')})

# 6. multi-GPU rows (scripts/scale_run.sh under gpurun --gpus N) next to the N=1 bench lines
scal = {}
for wl in ('detection', 'target'):
    rows = []
    for n in (1, 2, 4, 8):
        path = os.path.join(G, 'scale_%s_n%d.json' % (wl, n))
        if n == 1 and not os.path.exists(path):  # no N=1 line from the same library state: the bench line of this directory
            path = os.path.join(P, 'bench_%s_%s.json' % (wl, R))
        if not os.path.exists(path):
            continue
        txt = open(path).read().strip()
        d = json.loads(txt.splitlines()[-1] if txt.count('\n') == 0 or n > 1 else txt)
        rows.append({'n_gpus': d['n_gpus'], 'value': d['value'], 'ms_per_step': d['ms_per_step'], 'e2e': d['e2e']['value'],
                     'parity_check': d['parity_check']['result'], 'gather_check': d.get('gather_check'), 'scaling': d['scaling'],
                     'parallelism': d['config'].get('parallelism')})
    scal[wl] = rows
if any(len(v) > 1 for v in scal.values()):
    scal['note'] = 'N=1 rows are the bench lines of this directory; N>1: bench.py --gpus N --workload W under torchrun (scripts/scale_run.sh), all from the final library'
    json.dump(scal, open(os.path.join(P, 'scaling_%s.json' % R), 'w'), indent=1)

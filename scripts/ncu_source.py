"""Hottest CUDA source lines of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_source.py REPORT KERNEL_REGEX [TOP] [--nth N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k', 'regex:' + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fpath, hdr, data, func, seen_funcs = '', None, [], '', []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fpath = r[1].split('/')[-1]
    elif r[0] == 'Function Name':
        func = r[1]
        if func not in seen_funcs:
            seen_funcs.append(func)
    elif r[0] == 'Line No':
        hdr = r
    elif hdr and r[0] and len(r) > 8 and r[2] == '-' and func == seen_funcs[0]:
        try:
            samp = int(r[hdr.index('# Samples')] or 0)
            inst = int(r[hdr.index('Instructions Executed')] or 0)
        except ValueError:
            continue
        stalls = {}
        for name in ('stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_lg', 'stall_mio', 'stall_math', 'stall_wait',
                     'stall_branch_resolving', 'stall_not_selected', 'stall_no_inst', 'stall_membar'):
            try:
                v = int(r[hdr.index(name)] or 0)
            except ValueError:
                v = 0
            if v:
                stalls[name.replace('stall_', '')] = v
        data.append((samp, inst, fpath, r[0], r[1], stalls))
tot = sum(d[0] for d in data) or 1
print(seen_funcs[0][:90] if seen_funcs else '?', '| total samples', tot)
for samp, inst, f, ln, src, st in sorted(data, key=lambda d: -d[0])[:top]:
    ststr = ' '.join('%s=%d' % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print('%5.1f%% %9d inst %-13s L%-4s %-80s %s' % (100.0 * samp / tot, inst, f[:13], ln, src.strip()[:80], ststr))

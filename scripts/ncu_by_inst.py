"""Source lines of one kernel in an .ncu-rep ranked by executed warp instructions.
usage: ncu_by_inst.py REPORT KERNEL_REGEX [TOP]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k', 'regex:' + kern],
                     capture_output=True, text=True).stdout
hdr, func, first, agg = None, '', None, {}
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == 'Function Name':
        func = r[1]
        first = first or func
    elif r[0] == 'Line No':
        hdr = r
    elif hdr and r[0] and len(r) > 8 and r[2] == '-' and func == first:
        try:
            ln, inst, samp = int(r[0]), int(r[hdr.index('Instructions Executed')] or 0), int(r[hdr.index('# Samples')] or 0)
        except ValueError:
            continue
        if ln not in agg or agg[ln][0] < inst:
            agg[ln] = (inst, samp, r[1])
tot = sum(v[0] for v in agg.values()) or 1
print(first, '| warp instructions', tot)
for ln, (inst, samp, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%6.2f%% %9d inst %5d samples L%-5d %s' % (100 * inst / tot, inst, samp, ln, src.strip()[:100]))

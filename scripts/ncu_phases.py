"""Warp instructions and stall samples of one kernel in an .ncu-rep, binned by the SASS address ranges between the
kernel's %globaltimer reads (the DSPMB_STAMP / TraceScope markers), i.e. per phase in code-layout order.
usage: ncu_phases.py REPORT KERNEL_REGEX"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '-k', 'regex:' + kern],
                     capture_output=True, text=True).stdout
hdr, seen, bins, cur = None, 0, [], [0, 0, 0]
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == 'Kernel Name':
        seen += 1
        if seen > 1:
            break
    elif r[0] == 'Address':
        hdr = r
    elif hdr and r[0].startswith('0x'):
        inst = int(r[hdr.index('Instructions Executed')] or 0)
        samp = int(r[hdr.index('# Samples')] or 0)
        if 'GLOBALTIMER' in r[1]:
            bins.append(cur)
            cur = [0, 0, 0]
        cur[0] += inst
        cur[1] += samp
        cur[2] += 1
bins.append(cur)
ti, ts = sum(b[0] for b in bins) or 1, sum(b[1] for b in bins) or 1
print('total warp instructions %d, samples %d' % (ti, ts))
for i, b in enumerate(bins):
    print('segment %2d: %9d inst (%5.1f%%)  %6d samples (%5.1f%%)  %5d SASS lines' % (i, b[0], 100 * b[0] / ti, b[1], 100 * b[1] / ts, b[2]))

"""bench.py with tuning knobs from the environment: TUNE<k>=<v>."""
import os, sys
sys.path.insert(0, '.')
from dspnet_b200 import _lib
for k in range(32):
    v = os.environ.get('TUNE%d' % k)
    if v is not None:
        _lib.lib().dspmb_set_tuning(k, int(v))
import bench
sys.argv = ['bench.py'] + sys.argv[1:]
bench.main()

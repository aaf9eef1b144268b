import sys, os
sys.path.insert(0, '/root/repo')
from machupx_b200 import _native
if os.environ.get("LLS_LIB"): _native.LIB_PATH = os.path.join(os.path.dirname(_native.LIB_PATH), os.environ["LLS_LIB"])
sys.argv = ["lu_bench.py"] + sys.argv[1:]
exec(open('/root/repo/scripts/lu_bench.py').read())

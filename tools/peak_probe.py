import ctypes, sys
sys.path.insert(0, '/root/repo')
from haghighatshoarmuir2024_b200 import _native as N
lib = N.lib()
for v in range(4):
    x = ctypes.c_double()
    N.check(lib.micloc_fp32_peak(0, v, ctypes.byref(x)))
    print("variant", v, round(x.value, 2), "TFLOP/s")

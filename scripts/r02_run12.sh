ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum --clock-control none -k regex:k_eval_few -s 200 -c 3 --csv python scripts/oneq2.py 2>&1 | grep -i "k_eval_few" | awk -F'","' '{print $(NF-2), $(NF-1), $(NF)}'
python - <<'PY'
import ctypes
cuda = ctypes.CDLL("libcudart.so")
v = ctypes.c_int()
for name, attr in (("l2CacheSize", 38), ("maxPersistingL2CacheSize", 108), ("maxAccessPolicyWindowSize", 109)):
    cuda.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0); print(name, v.value)
PY

python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
b() { python bench.py --steps 20 --warmup 3 --no-cpu 2>>gpurun_out/err.txt | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$1',round(d['ms_per_step'],3),round(d['value']/1e6,1),round(d['e2e']['value']/1e6,1))"; }
b mb4
ncu --set full --clock-control none --import-source on -k regex:realign_kernel -s 2 -c 1 -o gpurun_out/r02_v4_realign python bench.py --steps 1 --warmup 2 --no-cpu > gpurun_out/ncu.log 2>&1
cp scratch_libs/libindelgpu_mb3.so indelminer_b200/libindelgpu.so
b mb3

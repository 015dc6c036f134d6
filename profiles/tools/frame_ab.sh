# A/B of two builds of the library on ONE box: bench.py (device-resident value only) with the in-tree build and with
# scratch/liblidar_b200_base.so, alternating.
L=lidar_ai_recommendation_software_b200/liblidar_b200.so
cp $L /tmp/new.so
run() { python bench.py --steps 10 --warmup 3 --extras "" --no-cpu --e2e-steps 12 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step']/512*1000,2), 'us/frame')"; }
for i in 1 2 3; do cp /tmp/new.so $L; run new; cp scratch/liblidar_b200_base.so $L; run base; done
cp /tmp/new.so $L

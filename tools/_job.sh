set -x
timeout 600 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_recommend.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --only lloyd --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])['lloyd']
print('lloyd ms', d['ms_per_step'], 'kmeans it', d['kmeans_iteration_ms'], d['kmeans_kernel_ms'])"
CRX_LLOYD_REFINE_UNSORTED=1 timeout 600 python bench.py --only lloyd --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])['lloyd']
print('unsorted: lloyd ms', d['ms_per_step'], 'kmeans it', d['kmeans_iteration_ms'])"
timeout 300 python tools/p2_probe.py 2>&1 | grep -E "wall|rec_finalize|tc_topp"

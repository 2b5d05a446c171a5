# A/B runs of the C3 step with parts of the fused backward kernels switched off (MMS_NVCC_EXTRA=-DMMS_BWD_PROBES build;
# the results are WRONG, only the times mean something).  All variants in ONE gpurun call: boxes differ by ~5 %.
run() { echo "== $*"; env "$@" python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels_ms_per_step']
print({n.split('<')[-1].strip('>'):v for n,v in k.items() if 'bwd_fused' in n or 'dm_kernel' in n or 'fwd' in n}, d['ms_per_step'])"; }
for v in "$@"; do run MMS_BWD_DEBUG=$v; done

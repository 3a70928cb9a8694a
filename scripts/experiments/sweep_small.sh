for cfg in "64 768" "64 256" "64 384" "128 256" "32 256" "128 512"; do
  set -- $cfg
  ML_FLSA_CHUNK_SMALL=$1 ML_FLSA_WARM_SMALL=$2 timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 3 > gpurun_out/r2i_sweep_$1_$2.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_sweep_$1_$2.json").read().strip().splitlines()[-1])
k={x['kernel']:x['ms_per_step'] for x in d['kernels']}
print("small chunk $1 warm $2: step %.3f ms spec %.3f chain %.3f"%(d['ms_per_step'],k.get('flsa_spec_kernel',0),k.get('flsa_chain_kernel',0)))
PY
done

run() {
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 3 > gpurun_out/r2i_s.json 2>/dev/null
  python - "$1" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r2i_s.json").read().strip().splitlines()[-1])
k={x['kernel']:x['ms_per_step'] for x in d['kernels']}
print("%s: step %.3f ms spec %.3f chain %.3f"%(sys.argv[1],d['ms_per_step'],k.get('flsa_spec_kernel',0),k.get('flsa_chain_kernel',0)))
PY
}
ML_FLSA_CHUNK_SMALL=64 ML_FLSA_WARM_SMALL=1024 run "small 64/1024"
ML_FLSA_CHUNK_SMALL=64 ML_FLSA_WARM_SMALL=1536 run "small 64/1536"
ML_FLSA_CHUNK_SMALL=32 ML_FLSA_WARM_SMALL=768 run "small 32/768"
ML_FLSA_CHUNK_BIG=384 ML_FLSA_WARM_BIG=640 run "big 384/640"
ML_FLSA_CHUNK_BIG=512 ML_FLSA_WARM_BIG=640 run "big 512/640"
ML_FLSA_CHUNK_BIG=320 ML_FLSA_WARM_BIG=512 run "big 320/512"

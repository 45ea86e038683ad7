# development aid: lock-step geometry sweep (warps per block x barrier mask)
for cfg in "8 22" "8 4" "10 22" "13 22"; do
  set -- $cfg
  echo -n "warps=$1 mask=$2: "
  GRS_STEP_WARPS=$1 GRS_LS_MASK=$2 python bench.py --steps 30 --warmup 10 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), round(d['ms_per_step'],2), round(d['roofline']['kernel_ms'],2))"
done

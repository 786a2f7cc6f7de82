for v in "VTC_PAIR=1" "VTC_PAIR=0" "VTC_CLUSTER=1" "VTC_PDL=0" "VTC_DBG_SKIP_EPILOGUE=1"; do
  echo "== $v"; env $v timeout 120 python scripts/tc_prof.py --d 256 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k!='cfg'})"
done
for v in "VTC_PAIR=1" "VTC_PAIR=0"; do
  echo "== bench $v"; env $v timeout 200 python bench.py --steps 10 --d 256 --no-cpu-baseline --no-e2e --no-extra | python scripts/show_bench.py /dev/stdin
done

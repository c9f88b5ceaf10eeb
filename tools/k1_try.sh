for v in "" ldcs "" ldcs; do
  if [ -n "$v" ]; then export AG_LIB=aprilgrid-rs_b200/lib/variants/libag_$v.so; else unset AG_LIB; fi
  echo "== variant '$v'"
  python bench.py --no-cpu --no-e2e --no-extras --steps 80 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('detect', d['value'], d['roofline']['frac'])"
done

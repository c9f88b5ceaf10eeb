p() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-28s' % sys.argv[1], round(d['value']), round(d['ms_per_step'],2), {k: round(v,1) for k,v in d['stage_ms_per_step'].items()}, d['tags_per_frame'])" "$1"; }
for cfg in ${SWEEP:-"4 512" "4 256" "2 512" "8 512" "1 512"}; do set -- $cfg
python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --board-warps $1 --chunk $2 | p "stream w$1 c$2"
done
python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --board-warps 4 --chunk 512 --sync-calls | p "sync w4 c512"

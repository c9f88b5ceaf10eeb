# streaming throughput against the pipeline chunk size and the warps per frame of the board kernel
for c in 256 384 512 768 1024; do for w in 2 4; do python bench.py --steps 24 --warmup 5 --no-e2e --no-cpu --chunk $c --board-warps $w 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk $c warps $w', round(d['value']), round(d['ms_per_step'],2))"; done; done

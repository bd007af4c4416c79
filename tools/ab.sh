run() { python bench.py --steps 6 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],2))"; }
DFB200_FUSE_RES=0 DFB200_FUSE_GG=0 DFB200_WGRAD7_DB=0 run base
DFB200_FUSE_RES=1 DFB200_FUSE_GG=0 DFB200_WGRAD7_DB=0 run res
DFB200_FUSE_RES=0 DFB200_FUSE_GG=1 DFB200_WGRAD7_DB=0 run gg
DFB200_FUSE_RES=0 DFB200_FUSE_GG=0 DFB200_WGRAD7_DB=1 run db7

run() { echo "== $1"; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --train-only 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        t=json.loads(l)['train']; a=t['allreduce']; print(round(t['ms_per_step'],3), 'nocomm', round(a['ms_per_step_without_allreduce'],3), 'alone', round(a['ms_alone'],3), 'exposed', round(a['exposed_ms'],3), 'buckets', a['buckets'])
"; }
run default X=1
run maxctas8 NCCL_MAX_CTAS=8
run maxctas4 NCCL_MAX_CTAS=4
run onebucket UNIPOSE_B200_BUCKET_BYTES=2000000000
run onebucket_ctas8 UNIPOSE_B200_BUCKET_BYTES=2000000000 NCCL_MAX_CTAS=8

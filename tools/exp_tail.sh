# train benches with the CUDA-graph replay on / off
for w in reddit wikipedia; do
for v in "TIGER_TRAIN_EAGER=0" "TIGER_TRAIN_EAGER=1"; do
env $v python bench.py --workload $w --mode train --steps 200 --warmup 20 --cpu-batches 10 2>gpurun_out/exp_train.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train $w $v',d['ms_per_step'],d['e2e']['value'],d.get('parity_checked'), d['train'].get('host_issue_ms_per_step'), d['train'].get('cuda_graph'), d['train'].get('mean_contrast_loss'), d['train'].get('mean_mutual_loss'))" || tail -5 gpurun_out/exp_train.err
done; done

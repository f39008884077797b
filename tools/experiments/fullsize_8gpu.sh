# BASELINE configs #4 and #3 at full replication counts on 8 GPUs (tools/showcase.py under torchrun)
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
$T tools/showcase.py --config 4 --chunk 262144 2> gpurun_out/full8_4.err | tail -1 > gpurun_out/full8_r02u.jsonl
$T tools/showcase.py --config 3 --chunk 50000 2> gpurun_out/full8_3.err | tail -1 >> gpurun_out/full8_r02u.jsonl
cat gpurun_out/full8_r02u.jsonl

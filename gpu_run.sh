mkdir -p gpurun_out
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
tail -2 gpurun_out/scale_$N.err | cut -c1-300; cut -c1-400 gpurun_out/scale_$N.json

N=$1; STEPS=${2:-100}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r2i_multi$N.log 2>&1; tail -4 gpurun_out/r2i_multi$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup 10 > gpurun_out/r2i_bench$N.log 2> gpurun_out/r2i_bench$N.err; tail -c 600 gpurun_out/r2i_bench$N.log; tail -3 gpurun_out/r2i_bench$N.err

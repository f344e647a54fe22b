T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_b_n8.json 2> gpurun_out/r02_b_n8.err
$T --master-port 29542 bench.py --gpus 8 --materials 20 --steps 21 --warmup 3 > gpurun_out/r02_b_n8_mat20.json 2> gpurun_out/r02_b_n8_mat20.err
$T --master-port 29543 bench.py --gpus 8 --size 512 --no-nce --steps 20 --warmup 3 > gpurun_out/r02_b_n8_512.json 2> gpurun_out/r02_b_n8_512.err
$T --master-port 29544 bench.py --gpus 8 --mode infer --size 1024 --batch 8 --steps 10 > gpurun_out/r02_infer_n8_b8.json 2> gpurun_out/r02_infer_n8_b8.err
$T --master-port 29545 bench.py --gpus 8 --mode infer --size 1024 --batch 32 --steps 5 > gpurun_out/r02_infer_n8_b32.json 2> gpurun_out/r02_infer_n8_b32.err
$T --master-port 29546 bench.py --gpus 8 --mode infer --size 1024 --batch 1 --steps 20 > gpurun_out/r02_infer_n8_b1.json 2> gpurun_out/r02_infer_n8_b1.err
tail -c 300 gpurun_out/r02_b_n8.json; tail -c 300 gpurun_out/r02_infer_n8_b32.json; tail -3 gpurun_out/r02_infer_n8_b32.err

N=${1:-8}
run() { env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tools/dma_rate.py 2>&1 | grep DMA_RATE; }
run MLV_DUMMY=1
run MLV_COPY_STREAMS=8
run MLV_COPY_CTAS=8 MLV_COPY_STREAMS=8
run MLV_COPY_CTAS=16 MLV_COPY_STREAMS=8
run MLV_COPY_CTAS=32 MLV_COPY_STREAMS=8

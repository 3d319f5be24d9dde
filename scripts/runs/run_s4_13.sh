for n in 8 4 2 1; do timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n scripts/exp_pcie_all.py 2>&1 | grep ranks; done

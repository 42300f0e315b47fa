#!/bin/bash
# torchrun --no-python wrapper: rank 0 runs its command under ncu (metrics of the kernels matching $NCU_KERNELS), the
# other ranks run it plainly.  NCU_OUT = csv log path.  Usage:
#   NCU_KERNELS='regex:k_ntt_shard' NCU_OUT=gpurun_out/x.csv python -m torch.distributed.run --no-python ... tools/ncu_rank0.sh python tools/sharded_proof_bench.py ...
if [ "${LOCAL_RANK:-0}" = "0" ]; then
    exec ncu --metrics "${NCU_METRICS:-gpu__time_duration.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum,dram__bytes_read.sum,dram__bytes_write.sum}" \
        --clock-control none -k "${NCU_KERNELS:-regex:k_ntt_shard}" -c "${NCU_COUNT:-20}" --csv --log-file "${NCU_OUT:-gpurun_out/ncu_rank0.csv}" "$@"
else
    exec "$@"
fi

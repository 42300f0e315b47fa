#!/usr/bin/env bash
# Evidence still owed for the domain-sharded NTT and the MSM-sharded proof (DESIGN.md sections 7 and 8), one gpurun
# call each:
#
#   gpurun --timeout 600 -- 'bash tools/gpu_evidence_sharded_ntt.sh one'
#   gpurun --gpus 2 --timeout 300 -- 'bash tools/gpu_evidence_sharded_ntt.sh two'
#   gpurun --gpus 8 --timeout 200 -- 'bash tools/gpu_evidence_sharded_ntt.sh eight'
#
# Everything lands in gpurun_out/; summaries to keep go to profiles/ by hand.
set -u
mkdir -p gpurun_out
case "${1:-one}" in
  one)
    # the four GPU tests added after round 1's last GPU call, then the whole sharded files
    timeout 400 python -m pytest tests/test_zz_gpu_unconfirmed.py tests/test_sharded.py tests/test_sharded_ntt.py tests/test_gpu_prove.py tests/test_gpu_host_mirror.py -m gpu -q 2>&1 | tail -8
    # launch list + full capture of one rank's four steps (world 8, 2^21: BASELINE config 5's per-rank work)
    timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/launches_ntt_shard.csv python tools/ntt_shard_time.py --logn 21 --worlds 8 \
        > gpurun_out/ntt_shard_time_under_ncu.log 2>&1
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_ntt_shard_(combine|split)' -c 4 \
        -o gpurun_out/ntt_shard_full -f python tools/ntt_shard_time.py --logn 21 --worlds 8 \
        > gpurun_out/ntt_shard_full.log 2>&1
    timeout 60 python tools/ntt_shard_time.py --logn 21 > gpurun_out/ntt_shard_time.jsonl 2> gpurun_out/ntt_shard_time.err
    # sanitizers over the simulated-world workload
    for tool in memcheck racecheck initcheck; do
      timeout 300 compute-sanitizer --tool $tool python tools/sanitize_sharded_ntt.py \
          > gpurun_out/sanitize_sharded_ntt_$tool.log 2>&1
      tail -2 gpurun_out/sanitize_sharded_ntt_$tool.log
    done
    ;;
  two)
    # NVLink side of the exchange kernels: ncu on rank 0 only would serialise the ranks, so time first ...
    timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29555 tools/ntt_shard_bench.py BN254:20 BLS12_381:21 \
        > gpurun_out/ntt_shard_bench_2gpu.jsonl 2> gpurun_out/ntt_shard_bench_2gpu.err
    cat gpurun_out/ntt_shard_bench_2gpu.jsonl
    # ... then the whole bench line with both sharded legs
    timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29556 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
    tail -c 1500 gpurun_out/bench_2gpu.json
    # ... and one proof over both GPUs (commit hook, MSMs sharded over the point set)
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29557 tools/sharded_proof_bench.py BN254:20 BLS12_381:20 \
        > gpurun_out/sharded_proof_2gpu.jsonl 2> gpurun_out/sharded_proof_2gpu.err
    cat gpurun_out/sharded_proof_2gpu.jsonl
    ;;
  eight)
    timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29555 tools/ntt_shard_bench.py BN254:20 BLS12_381:21 \
        > gpurun_out/ntt_shard_bench_8gpu.jsonl 2> gpurun_out/ntt_shard_bench_8gpu.err
    cat gpurun_out/ntt_shard_bench_8gpu.jsonl
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29557 tools/sharded_proof_bench.py BN254:20 BLS12_381:20 \
        > gpurun_out/sharded_proof_8gpu.jsonl 2> gpurun_out/sharded_proof_8gpu.err
    cat gpurun_out/sharded_proof_8gpu.jsonl
    ;;
esac

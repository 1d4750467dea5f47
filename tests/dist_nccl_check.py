"""Multi-GPU parity run (real kernels + NCCL all-to-all): launched with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist_nccl_check.py
Each rank runs the same DataStream programs (api_cases) and checks the gathered result against the oracle."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import torch
import torch.distributed as dist


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import api_cases as A
    from quokka_b200.df import QuokkaContext
    golden = os.path.join(HERE, "golden")
    for name in ("case_q1_sql", "case_q1_dict_api", "case_q3", "case_q5", "case_join_kinds", "case_asof",
                 "case_executor_protocol", "case_misc_ops", "case_scalar_aggs"):
        qc = QuokkaContext()
        qc.set_config("broadcast_rows", 100)        # shuffle (and Bloom-reduce) every join even at test sizes
        fn = getattr(A, name)
        if name in ("case_join_kinds", "case_asof", "case_executor_protocol"):
            fn(qc, golden)
        else:
            fn(qc)
        g = qc.last_graph
        if dist.get_rank() == 0:
            print(f"{name}: ok  (exchange calls {g.exchange.calls if g else 0}, via peer memory {g.exchange.peer_calls if g else 0}, bytes sent by rank 0 {g.exchange.bytes_sent if g else 0})", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if local == 0:
        print("DIST_NCCL_OK", flush=True)


if __name__ == "__main__":
    main()

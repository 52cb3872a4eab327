"""Worker of tests/test_gpu_dist.py: run under torchrun, one process per GPU.  Plays the same self-play and
tournament jobs as the single-process reference run of the test and leaves the files rank 0 wrote in the cwd."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    spec = json.loads(sys.argv[1])
    distributed = "RANK" in os.environ
    if distributed:
        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
    import training_pipeline as T
    os.makedirs("data/training_data", exist_ok=True)
    os.makedirs("data/tournament_results", exist_ok=True)
    fns = T.generate_Checkers_data(spec["selfplay"], spec["mcts"]).generate_data()
    tfn = T.tournament_Checkers(spec["tourney"], spec["tourney_mcts"]).start_tournament()
    if not distributed or dist.get_rank() == 0:
        json.dump({"data_fns": fns, "tourney_fn": tfn}, open("result.json", "w"))
    else:
        assert fns == [] and tfn is None
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

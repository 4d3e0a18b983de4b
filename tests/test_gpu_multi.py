"""N > 1 GPUs of one box (run with `gpurun --gpus 2`; skipped on a single-GPU box): ONE ensemble sharded over two ranks
(NCCL) must give, member for member, bit-identical results to the same ensemble on one GPU -- trained parameters, per-ROI
normative statistics, AUCs and per-subject deviations -- and every rank must hold the same gathered table."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank_main(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from multi_modal_normative_modeling_b200 import workloads
    from multi_modal_normative_modeling_b200.runner import EnsembleRunner
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    hw = workloads.build_host_workload(n_subjects=300, n_splits=3)
    sharded = EnsembleRunner(hw, 3, dev)                         # 3 folds x 4 modalities x 3 seeds = 36 members over 2 ranks
    sharded.train_epochs(3)
    gs = sharded.score()
    auc = sharded.fold_auc(gs)
    whole = EnsembleRunner(hw, 3, dev, rank=0, world=1)          # the same ensemble, all on this GPU, no collective
    whole.train_epochs(3)
    gw = whole.score()
    torch.cuda.synchronize()
    same_table = torch.equal(torch.nan_to_num(gs.table, nan=-1.0), torch.nan_to_num(gw.table, nan=-1.0))
    same_params = all(torch.equal(sharded.trainer.state_dict(k)[n], whole.trainer.state_dict(i)[n])
                      for k, i in enumerate(sharded.owned) for n in ("encoder_list.0.encoder_layers.0.weight",
                                                                      "decoder_list.0.decoder_mean_layer.bias"))
    torch.save({"owned": sharded.owned, "table": gs.table.cpu(), "same_table": same_table, "same_params": same_params,
                "auc": auc, "auc_whole": whole.fold_auc(gw)}, os.path.join(out_dir, f"rank{rank}.pt"))
    sharded.close(); whole.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_sharded_ensemble_equals_single_gpu_bit_for_bit(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert sorted(r0["owned"] + r1["owned"]) == list(range(36)) and len(r0["owned"]) == len(r1["owned"]) == 18
    assert r0["same_table"] and r1["same_table"] and r0["same_params"] and r1["same_params"]
    assert torch.equal(torch.nan_to_num(r0["table"], nan=-1.0), torch.nan_to_num(r1["table"], nan=-1.0))
    assert r0["auc"] == r1["auc"] == r0["auc_whole"] and len(r0["auc"]) == 9
    assert all(0.0 <= v <= 1.0 for v in r0["auc"].values())

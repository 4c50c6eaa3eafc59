"""Two-GPU test of the multi-GPU path (needs 2 visible GPUs: `gpurun --gpus 2`; skipped on a single-GPU box): utterance
sharding, the library's packed NCCL all-gather (gam_gather_hyps) and the empty-shard case, against the unsharded run."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import gigaam_b200 as gigaam
    from gigaam_b200 import dist as gdist, synthetic
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ok = True
        for name in ("v2_ctc", "v2_rnnt"):
            ck = synthetic.synthetic_checkpoint(name, seed=0, n_layers=2)
            model = gigaam.load_model(name, device=dev, checkpoint=ck)
            for batch in (5, 1, 4):                       # uneven split, an empty shard on rank 1, even split
                wav, wav_len = synthetic.synthetic_audio(batch, 2.0, seed=40 + batch)      # equal lengths: shards pad alike
                got = gdist.transcribe_sharded(model, wav, wav_len)
                enc, enc_len = model(wav.to(dev), wav_len.to(dev))
                want = model.decoding.decode(model.head, enc, enc_len)
                ok &= got == want and len(got) == batch
            ok &= int(model._get_engine().lib.gam_comm_nccl_version()) > 0
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_transcription_over_nccl_equals_unsharded():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}

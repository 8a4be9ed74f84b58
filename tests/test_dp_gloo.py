"""world_size-2 gloo tests (CPU) of the data-parallel host logic in textboost_b200/dp.py: row sharding, the
single flat-buffer all-reduce, and the identity the design rests on:

    N ranks x (B/N rows), SUM of per-rank-mean gradients, x 1/N  ==  1 rank x B rows (global mean)

The per-rank compute here is the oracle step on a tiny model (CPU); the CUDA path is covered by -m gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat_grads(te, out, V):
    parts = [out["grad_lora"][n].flatten() for n in sorted(out["grad_lora"])]
    parts.append(out["grad_rows"].flatten())
    return torch.cat(parts)


def _worker(rank, world, port, q, unet_adapter=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    torch.manual_seed(0)  # replicas must hold identical LoRA factors (set_seed, train_textboost.py:598-601)
    from textboost_b200 import dp
    from oracle import step_ref
    import test_oracle_cpu as T
    r, w, _ = dp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    sync = dp.GradSync()
    unet, te, te0, ccfg = T._tiny_step_models()
    if unet_adapter:  # --unet_params_to_train crossattn_kv: a second flat buffer, all-reduced like the first
        g = torch.Generator().manual_seed(7)
        for _, m in unet.add_cross_kv_lora(2):
            with torch.no_grad():
                m.lora_B["default"].weight.copy_(0.05 * torch.randn(m.lora_B["default"].weight.shape, generator=g))
    V = ccfg.vocab_size
    lat, noise, t, ids, pr = T._tiny_batch(4, V)
    sl = sync.shard_rows(4)
    assert (sl.start, sl.stop) == (2 * rank, 2 * rank + 2)
    out = step_ref.reference_step(unet, te, te0, lat[sl], noise[sl], t[sl], ids[sl], pr[sl], n_base=V)
    flat = _flat_grads(te, out, V)
    n_bytes = sync.payload_bytes(flat)
    sync.all_reduce_(flat)
    flat *= sync.inv_world

    def unet_flat(o):
        return torch.cat([o["grad_unet_lora"][n].flatten() for n in sorted(o["grad_unet_lora"])])
    if unet_adapter:
        flat_u = unet_flat(out)
        sync.all_reduce_(flat_u)
        flat_u *= sync.inv_world
    loss = out["loss"].clone()
    dist.all_reduce(loss)
    if rank == 0:
        full = step_ref.reference_step(unet, te, te0, lat, noise, t, ids, pr, n_base=V)
        if unet_adapter:
            flat, ref_flat = torch.cat([flat, flat_u]), torch.cat([_flat_grads(te, full, V), unet_flat(full)])
            assert flat_u.numel() > 0 and float(flat_u.abs().sum()) > 0
        else:
            ref_flat = _flat_grads(te, full, V)
        q.put((flat.tolist(), ref_flat.tolist(), (loss / world).item(), full["loss"].item(), n_bytes))
    dist.barrier()
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("unet_adapter", [False, True])
def test_two_rank_shard_allreduce_equals_global_batch(unet_adapter):
    """unet_adapter: with the UNet K/V LoRA (--unet_params_to_train crossattn_kv) a second flat gradient buffer goes
    through the same all-reduce (TextBoostTrainer.all_reduce), and the identity holds for it too."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, unet_adapter)) for r in range(2)]
    for p in procs:
        p.start()
    flat, ref, loss, loss_ref, n_bytes = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    flat, ref = torch.tensor(flat), torch.tensor(ref)
    torch.testing.assert_close(flat, ref, rtol=1e-4, atol=1e-7)
    assert abs(loss - loss_ref) < 1e-5 * abs(loss_ref)
    assert unet_adapter or n_bytes == flat.numel() * 4


def test_single_process_sync_is_identity():
    from textboost_b200 import dp
    s = dp.GradSync()
    g = torch.arange(8, dtype=torch.float32)
    assert s.world == 1 and s.inv_world == 1.0 and torch.equal(s.all_reduce_(g.clone()), g)
    assert s.shard_rows(8) == slice(0, 8)

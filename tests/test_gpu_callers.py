"""-m gpu: the reference's CALLERS replayed on the drop-in modules, plus the optimiser / patching / criterion variants
the shipped YAML can select.

The reference itself is not on the GPU box, so its call sequences are restated here line by line:
  * tools/train.py:200-211 (constructor keywords), :226-241 (optimiser selection), :245-248 (criterion), :280-298
    (training iteration with ``.item()`` / ``.cpu()`` reads);
  * models/mpp.py:77-134 -- the reference MPP forward body reaching INTO a SiT through its attribute surface
    (``to_patch_embedding[-1]``, ``cls_token``, ``pos_embedding``, ``dropout``, ``transformer(x)``), here driving the
    B200 ``SiT``.
Everything is compared with the fp32 oracle on identical weights, inputs and masks (bf16 tolerance 1e-2)."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch import nn, optim

from helpers import rel_l2
from oracle.sit_oracle import OracleMPP, OracleSiT

pytestmark = pytest.mark.gpu

import surface_vision_transformers_b200 as svit  # noqa: E402
from surface_vision_transformers_b200 import _lib  # noqa: E402
from surface_vision_transformers_b200._lib import check, ptr, vp  # noqa: E402

TOL = 1e-2
DEV = torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _fp32_oracle():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def worst_grad(ours, oracle):
    ref = dict(oracle.named_parameters())
    return max((rel_l2(p.grad, ref[n].grad), n) for n, p in ours.named_parameters() if p.grad is not None)


# ---------------------------------------------------------------------------------------------------------------
# tools/train.py on the drop-in
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("optimiser,use_l1loss", [("SGD", False), ("AdamW", False), ("Adam", True)])
def test_train_py_call_sequence(optimiser, use_l1loss):
    """tools/train.py:200-211, 226-248, 280-298 verbatim on the B200 SiT (left) and on the oracle (right): same YAML
    keywords, the script's own torch.optim optimisers (SGD momentum 0.9 is the YAML default, hparams.yml:50-61), the
    per-iteration host reads.  Losses of every iteration and the epoch MAE agree within the bf16 tolerance."""
    config = {"transformer": dict(dim=192, depth=3, heads=3, mlp_dim=768, pool="cls", num_classes=1, num_channels=4, dim_head=64,
                                  dropout=0.0, emb_dropout=0.0),
              "SGD": dict(weight_decay=0.0, momentum=0.9, nesterov=False), "Adam": dict(weight_decay=0.0),
              "AdamW": dict(weight_decay=0.0)}
    num_patches, num_vertices, LR, device = 80, 45, 1e-4, DEV
    torch.manual_seed(0)
    models = []
    for SiT in (svit.SiT, OracleSiT):
        model = SiT(dim=config['transformer']['dim'],                       # train.py:200-211
                    depth=config['transformer']['depth'],
                    heads=config['transformer']['heads'],
                    mlp_dim=config['transformer']['mlp_dim'],
                    pool=config['transformer']['pool'],
                    num_patches=num_patches,
                    num_classes=config['transformer']['num_classes'],
                    num_channels=config['transformer']['num_channels'],
                    num_vertices=num_vertices,
                    dim_head=config['transformer']['dim_head'],
                    dropout=config['transformer']['dropout'],
                    emb_dropout=config['transformer']['emb_dropout'])
        models.append(model)
    models[0].load_state_dict(models[1].state_dict())
    data = [(torch.randn(6, 4, num_patches, num_vertices), torch.rand(6) * 19 + 26) for _ in range(4)]
    histories = []
    for model in models:
        model.to(device)                                                      # train.py:226
        if optimiser == 'Adam':                                               # train.py:228-241
            optimizer = optim.Adam(model.parameters(), lr=LR, weight_decay=config['Adam']['weight_decay'])
        elif optimiser == 'SGD':
            optimizer = optim.SGD(model.parameters(), lr=LR, weight_decay=config['SGD']['weight_decay'],
                                  momentum=config['SGD']['momentum'], nesterov=config['SGD']['nesterov'])
        else:
            optimizer = optim.AdamW(model.parameters(), lr=LR, weight_decay=config['AdamW']['weight_decay'])
        criterion = nn.MSELoss(reduction='mean') if not use_l1loss else nn.L1Loss()     # train.py:245-248
        running_loss, targets_, preds_, losses = 0, [], [], []
        model.train()
        for i, d in enumerate(data):                                          # train.py:280-298
            inputs, targets = d[0].to(device), d[1].to(device)
            optimizer.zero_grad()
            outputs = model(inputs)
            loss = criterion(outputs.squeeze(), targets)
            loss.backward()
            optimizer.step()
            running_loss += loss.item()
            losses.append(loss.item())
            targets_.append(targets.cpu().numpy())
            preds_.append(outputs.reshape(-1).cpu().detach().numpy())
        mae_epoch = np.mean(np.abs(np.concatenate(targets_) - np.concatenate(preds_)))
        histories.append((losses, mae_epoch, optimizer.param_groups[0]['lr']))
    (l_new, mae_new, lr_new), (l_ref, mae_ref, lr_ref) = histories
    for a, b in zip(l_new, l_ref):
        assert abs(a - b) / abs(b) < TOL, (l_new, l_ref)
    assert abs(mae_new - mae_ref) / mae_ref < TOL and lr_new == lr_ref
    # validation call of train.py:311-337: eval + no_grad, batch size 1 (bs_val: 1)
    models[0].eval(); models[1].eval()
    with torch.no_grad():
        x1 = data[0][0][:1].to(device)
        assert rel_l2(models[0](x1), models[1](x1)) < 3 * TOL


# ---------------------------------------------------------------------------------------------------------------
# the reference's models/mpp.py forward body on a B200 SiT
# ---------------------------------------------------------------------------------------------------------------
def reference_mpp_forward(self_transformer, to_original, mask_token, batch, masks):
    """models/mpp.py:77-134 with the random draws replaced by the given masks (same substitution as the oracle)."""
    corrupted_sequence, bool_random_patch_prob, random_patches, bool_mask_replace = masks
    transformer = self_transformer
    batch = batch.permute(0, 2, 3, 1).reshape(batch.shape[0], batch.shape[2], -1)     # 'b c n v -> b n (v c)'  :81-82
    corrupted_batch = batch.clone().detach()                                            # :87
    if bool_random_patch_prob is not None:                                              # :90-107
        randomized_input = corrupted_batch[torch.arange(corrupted_batch.shape[0]).unsqueeze(-1), random_patches]
        corrupted_batch[bool_random_patch_prob] = randomized_input[bool_random_patch_prob]
    corrupted_batch[bool_mask_replace] = mask_token.to(corrupted_sequence.device)       # :112
    corrupted_batch = transformer.to_patch_embedding[-1](corrupted_batch)               # :115
    b, n, _ = corrupted_batch.shape
    cls_tokens = transformer.cls_token.expand(b, -1, -1)                                # :120  repeat '() n d -> b n d'
    corrupted_batch = torch.cat((cls_tokens, corrupted_batch), dim=1)                   # :121
    corrupted_batch += transformer.pos_embedding[:, :(n + 1)]                           # :124
    corrupted_batch = transformer.dropout(corrupted_batch)                              # :125
    batch_out = transformer.transformer(corrupted_batch)                                # :128
    batch_out = to_original(batch_out[:, 1:, :])                                        # :129
    mpp_loss = F.mse_loss(batch_out[corrupted_sequence], batch[corrupted_sequence])     # :132
    return mpp_loss, batch_out


def test_reference_mpp_body_drives_b200_sit_through_its_attributes():
    """The UNMODIFIED reference MPP module never calls SiT.forward: it uses the patch-embedding Linear, cls / pos
    parameters, the dropout module and ``transformer.transformer`` directly.  That sequence on a B200 SiT (eager torch
    for the embedding glue exactly as in the reference, the fused encoder for ``transformer(x)``, autograd through both)
    reproduces the oracle's loss, reconstruction and every gradient."""
    cfg = dict(dim=192, depth=3, heads=3, mlp_dim=768, num_patches=80, num_vertices=45)
    B, K = 4, 4 * 45
    torch.manual_seed(5)
    kw = dict(mask_prob=0.5, replace_prob=0.8, swap_prob=0.1, channels=4, num_vertices=45)
    oracle = OracleMPP(OracleSiT(**cfg), cfg["dim"], K, DEV, **kw).to(DEV)
    sit = svit.SiT(**cfg)
    sit.load_state_dict(oracle.transformer.state_dict())
    sit.to(DEV)
    to_original = nn.Linear(cfg["dim"], K).to(DEV)                                       # mpp.py:66-67
    to_original.load_state_dict(oracle.to_original.state_dict())
    mask_token = nn.Parameter(oracle.mask_token.detach().clone())                        # mpp.py:74
    x = torch.randn(B, 4, cfg["num_patches"], cfg["num_vertices"], device=DEV)
    from surface_vision_transformers_b200.mpp import draw_masks
    masks = draw_masks(B, cfg["num_patches"], K, DEV, 0.5, 0.8, 0.1)
    lo, oo = oracle(x, masks=masks)
    lo.backward()
    lm, om = reference_mpp_forward(sit, to_original, mask_token, x, masks)
    lm.backward()
    assert abs(lm.item() - lo.item()) / lo.item() < TOL
    assert rel_l2(om, oo) < TOL
    ref = dict(oracle.named_parameters())
    worst = (0.0, None)
    for n, p in sit.named_parameters():
        r = ref["transformer." + n]
        if r.grad is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        worst = max(worst, (rel_l2(p.grad, r.grad), n))
    worst = max(worst, (rel_l2(to_original.weight.grad, ref["to_original.weight"].grad), "to_original.weight"),
                (rel_l2(to_original.bias.grad, ref["to_original.bias"].grad), "to_original.bias"),
                (rel_l2(mask_token.grad, ref["mask_token"].grad), "mask_token"))
    print(f"reference-mpp-body on B200 SiT: worst gradient tensor {worst[1]} rel-L2 {worst[0]:.2e}")
    assert worst[0] < TOL, worst


# ---------------------------------------------------------------------------------------------------------------
# SGD (the YAML default optimiser) through the fused kernel
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("momentum,nesterov,wd,damp", [(0.9, False, 0.0, 0.0), (0.9, True, 1e-4, 0.0), (0.0, False, 1e-3, 0.0),
                                                      (0.8, False, 0.0, 0.1)])
def test_sgd_kernel_matches_torch(momentum, nesterov, wd, damp):
    """svit_sgd_step vs torch.optim.SGD (hparams.yml:57-60: momentum 0.9, nesterov False / True), 4 steps incl. the first
    (momentum buffer = gradient), with a gradient scale (the 1/world factor of the data-parallel path)."""
    lib = _lib.load()
    torch.manual_seed(0)
    n = 100_003
    p0 = torch.randn(n, device=DEV)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([ref_p], lr=0.05, momentum=momentum, nesterov=nesterov, weight_decay=wd, dampening=damp)
    p = p0.clone()
    mom = torch.zeros(n, device=DEV)
    for step in range(4):
        g = torch.randn(n, device=DEV)
        ref_p.grad = g * 0.5
        opt.step()
        check(lib.svit_sgd_step(ptr(p), ptr(g), ptr(mom), n, 0.05, momentum, damp, wd, 1 if nesterov else 0,
                                1 if step == 0 else 0, 0.5, vp(torch.cuda.current_stream().cuda_stream)), "sgd")
    torch.cuda.synchronize()
    assert rel_l2(p, ref_p.detach()) < 1e-6
    if momentum != 0:
        assert rel_l2(mom, opt.state[ref_p]["momentum_buffer"]) < 1e-6


def test_fused_sgd_tracks_oracle_and_reloads_its_state():
    """FusedSGD (one kernel over the flat buffer) vs torch.optim.SGD on the oracle over 3 steps; then state_dict() ->
    a fresh optimizer -> load_state_dict(): the restored momentum is USED (the next step equals an uninterrupted run)."""
    cfg = dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=12, num_vertices=10)
    torch.manual_seed(0)
    oracle = OracleSiT(**cfg).to(DEV)
    model = svit.SiT(**cfg)
    model.load_state_dict(oracle.state_dict())
    model.to(DEV)
    start = copy.deepcopy(oracle)
    kw = dict(lr=1e-3, momentum=0.9, nesterov=True, weight_decay=1e-4)
    o1 = torch.optim.SGD(oracle.parameters(), **kw)
    o2 = svit.FusedSGD(model.parameters(), **kw)
    lib = _lib.load()
    batches = [(torch.randn(5, 4, 12, 10, device=DEV), torch.rand(5, device=DEV) * 19 + 26) for _ in range(4)]

    def step(m, o, xy):
        o.zero_grad()
        loss = F.mse_loss(m(xy[0]).squeeze(), xy[1])
        loss.backward()
        o.step()
        return loss.item()

    for xy in batches[:3]:
        l1, l2 = step(oracle, o1, xy), step(model, o2, xy)
        assert abs(l1 - l2) / l1 < TOL
    moved = sum(((p - q) ** 2).sum() for p, q in zip(oracle.parameters(), start.parameters())).sqrt()
    diff = sum(((p - q) ** 2).sum() for p, q in zip(oracle.parameters(), model.parameters())).sqrt()
    assert moved > 0 and (diff / moved).item() < 2 * TOL, (diff / moved).item()
    # resume: clone the model, reload the optimizer state into a fresh FusedSGD, take one more step on both
    twin = svit.SiT(**cfg)
    twin.load_state_dict(model.state_dict())
    twin.to(DEV)
    o3 = svit.FusedSGD(twin.parameters(), **kw)
    o3.load_state_dict(copy.deepcopy(o2.state_dict()))
    n0 = lib.svit_launch_count()
    step(model, o2, batches[3])
    step(twin, o3, batches[3])
    assert lib.svit_launch_count() > n0
    assert rel_l2(twin._flat, model._flat) < 1e-6          # identical arithmetic: the momentum survived the reload


def test_fused_adamw_updates_mpp_parameters_with_the_kernel():
    """to_original.* and mask_token of the MPP module live in a flat buffer of their own and name it as owner: a
    FusedAdamW over ssl.parameters() updates them with one svit_adamw_step launch per module (no eager arithmetic),
    and the result tracks torch.optim.AdamW on the oracle."""
    cfg = dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=20, num_vertices=15)
    K = 60
    kw = dict(mask_prob=0.5, replace_prob=0.8, swap_prob=0.1, channels=4, num_vertices=15)
    torch.manual_seed(4)
    oracle = OracleMPP(OracleSiT(**cfg), cfg["dim"], K, DEV, **kw).to(DEV)
    ssl = svit.masked_patch_pretraining(transformer=svit.SiT(**cfg), dim_in=cfg["dim"], dim_out=K, device=DEV, **kw)
    ssl.load_state_dict(oracle.state_dict())
    ssl.to(DEV)
    for p in (ssl.to_original.weight, ssl.to_original.bias, ssl.mask_token):
        assert p._svit_owner() is ssl
    o1 = torch.optim.AdamW(oracle.parameters(), lr=1e-3, weight_decay=0.0)
    o2 = svit.FusedAdamW(ssl.parameters(), lr=1e-3, weight_decay=0.0)
    def no_eager(group, params):          # the eager per-tensor branch must not see any parameter of this job
        assert not [p for p in params if p.grad is not None], "a parameter fell back to eager arithmetic"
    o2._generic_adam = no_eager
    from surface_vision_transformers_b200.mpp import draw_masks
    w0 = oracle.to_original.weight.detach().clone()
    for it in range(3):
        x = torch.randn(4, 4, 20, 15, device=DEV)
        masks = draw_masks(4, 20, K, DEV, 0.5, 0.8, 0.1)
        o1.zero_grad(); o2.zero_grad()
        lo, _ = oracle(x, masks=masks); lo.backward(); o1.step()
        lm, _ = ssl(x, masks=masks); lm.backward(); o2.step()
        assert abs(lm.item() - lo.item()) / lo.item() < TOL
    # every MPP parameter carries fused state (views of the flat moment buffers), none took the eager branch
    st = o2.state[ssl.to_original.weight]
    assert st["exp_avg"].data_ptr() != 0 and float(st["step"]) == 3.0
    assert st["exp_avg"].untyped_storage().data_ptr() == o2.state[ssl.mask_token]["exp_avg"].untyped_storage().data_ptr()
    moved = (oracle.to_original.weight - w0).norm()
    assert moved > 0 and ((ssl.to_original.weight - oracle.to_original.weight).norm() / moved).item() < 0.1


# ---------------------------------------------------------------------------------------------------------------
# sub_ico_0 patching, L1 criterion, full batch
# ---------------------------------------------------------------------------------------------------------------
def test_sub_ico_0_patching_vs_oracle():
    """hparams.yml:71-73: sub_ico_0 = 20 patches x 2145 vertices (K = 8580, T = 21) -- also the constructor defaults of
    models/sit.py:31-35.  Forward, encoder output and every gradient against the oracle."""
    cfg = dict(dim=192, depth=2, heads=3, mlp_dim=768)          # num_patches=20, num_vertices=2145 are the defaults
    B = 4
    torch.manual_seed(0)
    oracle = OracleSiT(**cfg).to(DEV)
    model = svit.SiT(**cfg)
    assert model.num_patches == 20 and model.num_vertices == 2145
    model.load_state_dict(oracle.state_dict())
    model.to(DEV)
    x = torch.randn(B, 4, 20, 2145, device=DEV)
    y = torch.rand(B, device=DEV) * 19 + 26
    out_o = oracle(x); F.mse_loss(out_o.squeeze(), y).backward()
    out_m = model(x); F.mse_loss(out_m.squeeze(), y).backward()
    assert rel_l2(out_m, out_o) < 3 * TOL
    w = worst_grad(model, oracle)
    print(f"sub_ico_0: worst gradient tensor {w[1]} rel-L2 {w[0]:.2e}")
    assert w[0] < TOL, w
    with torch.no_grad():
        model.eval(); oracle.eval()
        assert rel_l2(model(x[:1]), oracle(x[:1])) < 3 * TOL


def test_fit_with_l1_criterion(tmp_path):
    """train.py:247-248 (l1loss: True) through fit(): per-epoch L1 loss / MAE of a reference-style loop on the oracle."""
    cfg = dict(dim=128, depth=2, heads=2, mlp_dim=256, num_patches=12, num_vertices=10)
    rng = np.random.default_rng(3)
    for split, n in (("train", 20), ("validation", 6)):
        np.save(tmp_path / f"{split}_data.npy", rng.standard_normal((n, 4, 12, 10)))
        np.save(tmp_path / f"{split}_labels.npy", rng.uniform(26, 45, size=n))
    train = svit.PatchedNpyDataset(str(tmp_path), "train")
    val = svit.PatchedNpyDataset(str(tmp_path), "validation")
    torch.manual_seed(0)
    oracle = OracleSiT(**cfg).to(DEV)
    model = svit.SiT(**cfg)
    model.load_state_dict(oracle.state_dict())
    model.to(DEV)
    res = svit.fit(model, svit.FusedSGD(model.parameters(), lr=1e-3, momentum=0.9), train, val, epochs=2, batch_size=8,
                   device=DEV, l1loss=True, seed=5)
    opt = torch.optim.SGD(oracle.parameters(), lr=1e-3, momentum=0.9)
    crit = nn.L1Loss()
    gen = torch.Generator()
    for epoch in range(2):
        gen.manual_seed(5 + epoch)
        oracle.train()
        run, nb = 0.0, 0
        for x, y in train.batches(8, shuffle=True, generator=gen):
            x, y = x.to(DEV), y.to(DEV)
            opt.zero_grad()
            loss = crit(oracle(x).squeeze(), y)
            loss.backward()
            opt.step()
            run += loss.item(); nb += 1
        ref = run / nb
        assert abs(res["history"]["train_loss"][epoch] - ref) / ref < 3 * TOL, (epoch, res["history"]["train_loss"], ref)
    oracle.eval()
    with torch.no_grad():
        vp_ = torch.cat([oracle(x.to(DEV)).reshape(-1).cpu() for x, _ in val.batches(8)])
    ref_vmae = (val.labels - vp_).abs().mean().item()
    assert abs(res["history"]["val_mae"][-1][1] - ref_vmae) / ref_vmae < 3 * TOL


def test_full_batch_256_two_layer_slice_vs_oracle():
    """BASELINE configs[1] at its full per-GPU batch (256) on a 2-block slice of SiT-small: encoder output and every
    gradient against the fp32 oracle (the 12-block comparison runs at batch 6, test_sit_vs_oracle_full_configs)."""
    cfg = dict(dim=384, depth=2, heads=6, mlp_dim=1536, num_patches=320, num_vertices=153)
    B = 256
    torch.manual_seed(0)
    oracle = OracleSiT(**cfg).to(DEV)
    model = svit.SiT(**cfg)
    model.load_state_dict(oracle.state_dict())
    model.to(DEV)
    x = torch.randn(B, 4, 320, 153, device=DEV)
    y = torch.rand(B, device=DEV) * 19 + 26
    out_o = oracle(x); F.mse_loss(out_o.squeeze(), y).backward()
    out_m = model(x); F.mse_loss(out_m.squeeze(), y).backward()
    w = worst_grad(model, oracle)
    print(f"B=256 two-layer slice: worst gradient tensor {w[1]} rel-L2 {w[0]:.2e}")
    assert w[0] < TOL, w
    assert rel_l2(out_m, out_o) < 3 * TOL
    with torch.no_grad():
        xe = oracle.to_patch_embedding(x[:64])
        xe = torch.cat((oracle.cls_token.expand(64, -1, -1), xe), 1) + oracle.pos_embedding
        assert rel_l2(model.transformer(xe), oracle.transformer(xe)) < TOL


def test_training_from_raw_meshes_vs_oracle():
    """SURVEY 8(f)-1 with autograd: forward_mesh() gathers (tools/preprocessing.py:79-84) and z-scores (:72) the raw ico-6
    meshes inside the kernel that packs the patch-embedding operand, and trains from there.  Reference side: the numpy
    restatement of the reference's preprocessing produces the patched array, the fp32 oracle trains on it.  Loss,
    prediction and every gradient (the patch-embedding weight in particular) agree within the bf16 tolerance."""
    from oracle import gather_oracle
    cfg = dict(dim=192, depth=2, heads=3, mlp_dim=768, num_patches=320, num_vertices=153)
    B = 3
    torch.manual_seed(0)
    rs = np.random.RandomState(11)
    mesh = (rs.standard_normal((B, 4, 40962)) * 2 + 1).astype(np.float32)
    means = np.array([1.15, 0.037, 1.0, 0.07], dtype=np.float32)
    stds = np.array([0.41, 0.19, 0.39, 4.05], dtype=np.float32)
    table = svit.load_index_table(2)
    normalised = (mesh - means.reshape(1, 4, 1)) / stds.reshape(1, 4, 1)
    patched = torch.from_numpy(gather_oracle.gather_patches(normalised, table.numpy()).astype(np.float32))
    assert tuple(patched.shape) == (B, 4, 320, 153)
    oracle = OracleSiT(**cfg).to(DEV)
    model = svit.SiT(**cfg)
    model.load_state_dict(oracle.state_dict())
    model.to(DEV)
    y = torch.rand(B, device=DEV) * 19 + 26
    out_o = oracle(patched.to(DEV)); lo = F.mse_loss(out_o.squeeze(), y); lo.backward()
    out_m = model.forward_mesh(torch.from_numpy(mesh).to(DEV), table.to(DEV), torch.from_numpy(means).to(DEV),
                               torch.from_numpy(stds).to(DEV))
    lm = F.mse_loss(out_m.squeeze(), y); lm.backward()
    assert rel_l2(out_m, out_o) < 3 * TOL and abs(lm.item() - lo.item()) / lo.item() < 3 * TOL
    w = worst_grad(model, oracle)
    print(f"training from raw meshes: worst gradient tensor {w[1]} rel-L2 {w[0]:.2e}")
    assert w[0] < TOL, w


def test_bf16_staged_batches_give_bit_identical_results(tmp_path):
    """svit_forward_ex(input_bf16=1): a batch rounded to bf16 on the host (PatchedNpyDataset(stage_dtype=bfloat16), half
    the host-to-device bytes of tools/train.py:281-283) gives the SAME bits as its fp32 copy -- prediction, loss and every
    gradient -- because the packing kernel rounds fp32 input to bf16 with the same round-to-nearest-even."""
    import numpy as np
    cfg = dict(dim=192, depth=2, heads=3, mlp_dim=768, num_patches=80, num_vertices=45)
    torch.manual_seed(3)
    model = svit.SiT(**cfg).to(DEV)
    x = torch.randn(6, 4, 80, 45, device=DEV)
    y = torch.rand(6, device=DEV) * 19 + 26
    xb = x.bfloat16()
    res = []
    for inp in (xb, xb.float(), x):
        model.zero_grad(set_to_none=True)
        out = model(inp)
        torch.nn.functional.mse_loss(out.squeeze(), y).backward()
        res.append((out.detach().clone(), model.to_patch_embedding[1].weight.grad.clone(), model.pos_embedding.grad.clone()))
    assert torch.equal(res[0][0], res[1][0])                      # bf16 batch == its fp32 copy, bit for bit (the forward has no atomics)
    assert torch.equal(res[0][0], res[2][0])                      # ... and == the original fp32 batch (same rounding in the kernel)
    for k in (1, 2):                                              # gradients: equal up to the fp32 atomics order of the weight-gradient kernels
        assert rel_l2(res[0][k], res[1][k]) < 1e-6 and rel_l2(res[0][k], res[2][k]) < 1e-6
    with torch.no_grad():
        assert torch.equal(model.eval()(xb), model(xb.float()))
    # the reader hands out pinned bf16 batches
    np.save(tmp_path / "train_data.npy", x.double().cpu().numpy())
    np.save(tmp_path / "train_labels.npy", y.double().cpu().numpy())
    ds = svit.PatchedNpyDataset(str(tmp_path), "train", stage_dtype=torch.bfloat16)
    bx, by = next(iter(ds.batches(6)))
    assert bx.dtype == torch.bfloat16 and by.dtype == torch.float32 and torch.equal(bx.to(DEV), xb)
    # fp32 check mode has no bf16 operand: the module converts
    model.set_check_mode(True)
    with torch.no_grad():
        assert torch.equal(model(xb), model(xb.float()))


@pytest.mark.parametrize("l1", [False, True])
def test_fused_regression_loss_matches_torch_criterion(l1):
    """svit.regression_loss == nn.MSELoss(reduction='mean') / nn.L1Loss() on outputs.squeeze() (train.py:245-248, 288):
    value and gradient, incl. an exact zero residual (sign(0) = 0 for L1) and an upstream scale."""
    torch.manual_seed(0)
    out = (torch.rand(37, 1, device=DEV) * 19 + 26).requires_grad_(True)
    tgt = torch.rand(37, device=DEV) * 19 + 26
    with torch.no_grad():
        out[5, 0] = tgt[5]
    ref_out = out.detach().clone().requires_grad_(True)
    crit = nn.L1Loss() if l1 else nn.MSELoss(reduction="mean")
    ref = crit(ref_out.squeeze(), tgt)
    (ref * 3.0).backward()
    got = svit.regression_loss(out, tgt, l1loss=l1)
    (got * 3.0).backward()
    assert got.dim() == 0 and abs(got.item() - ref.item()) / ref.item() < 1e-6
    assert torch.allclose(out.grad, ref_out.grad, rtol=1e-6, atol=1e-9)
    assert float(out.grad[5, 0]) == 0.0

"""CPU: the reference's OWN files, unchanged, on the MinkowskiEngine-shaped shim -- proven equal to the mirrors the
bench and the GPU tests run (missing item 1 / next 2c).

  * utils/models/minkunet_bev.py:302-399   MinkUNet34BEV.forward           == lidog_b200/lidog/model.py (bit-equal)
  * utils/collation/collation.py:274-325   CollateFNSingleSourceBEVMultiLevel (runs on ME.utils.SparseCollation)
  * utils/pipelines/trainer_lighting_2d.py:141-293  PLTTrainer2D.training_step + configure_optimizers
                                                                           == lidog_b200/lidog/step.py (loss, gradients)
  * utils/losses/losses.py:56-97,129-187   DICELoss / SoftDICELoss         == lidog_b200/lidog/losses.py
  * utils/datasets/synth4d_bev.py:478-509  PC2ImgConverter.getBEVImageNew  == lidog_b200/lidog/step.bev_label_image
The shim under these tests is the CPU oracle's stand-in (there is no GPU here); the CUDA product implements the same
surface and is compared with the oracle in the -m gpu tests, so equality here carries the unchanged files onto the CUDA
path.  Skipped when /root/reference is absent (the GPU box)."""
import numpy as np
import pytest
import torch

from tests import refharness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="/root/reference is not mounted here")


def _scan(seed=3, r=8.0, shape="nuscenes"):
    from lidog_b200.lidog import synth
    pts, lab = synth.make_scan(seed, shape)
    keep = (np.abs(pts[:, 0]) < r) & (np.abs(pts[:, 1]) < r)
    return pts[keep], lab[keep]


def _dataset_item(pts, lab, me, bound, img, voxel=0.05):
    """What the reference dataset's __getitem__ hands to the collation (semantickitti_bev.py:232-252), built with the
    shim's sparse_quantize; `bev_labels` through the reference's own PC2ImgConverter."""
    feats = np.ones((len(pts), 1), np.float32)
    q, f, colab, vidx, inv = me.utils.sparse_quantize(pts, feats, labels=lab, ignore_label=-1, quantization_size=voxel,
                                                      return_index=True, return_inverse=True)
    return q, f, colab, vidx


def test_reference_model_file_runs_unchanged_and_equals_the_mirror():
    from oracle import me_cpu
    from oracle.me_cpu.bevfn import sparse2super as o_s2s
    from oracle import voxel as ov
    from lidog_b200.lidog.model import MinkUNet34BEV as Mirror
    pts, lab = _scan()
    q = ov.sparse_quantize(pts, quantization_size=0.05)
    coords = torch.from_numpy(ov.batched_coordinates([q]))
    with rh.reference(me_cpu):
        import utils.models.minkunet_bev as ref
        torch.manual_seed(0)
        r = ref.MinkUNet34BEV(1, 7, 3, mapping_bound_2d=8.0)
        m = Mirror(1, 7, ME=me_cpu, bev_fn=o_s2s, mapping_bound_2d=8.0)
        m.load_state_dict(r.state_dict())  # identical names and shapes, or this raises
        torch.set_num_threads(1)  # the reference's index_put_ overwrite is deterministic only single-threaded
        try:
            out_r, bev_r = r(me_cpu.SparseTensor(coordinates=coords, features=torch.ones(len(q), 1)), is_train=True)
            out_m, bev_m = m(me_cpu.SparseTensor(coordinates=coords, features=torch.ones(len(q), 1)), is_train=True)
            assert torch.equal(out_r.F, out_m.F)                      # bit-equal logits
            assert torch.equal(bev_r["block8"], bev_m["block8"])      # bit-equal BEV logits (sparse2super + Encoder2D)
            (out_r.F.square().mean() + bev_r["block8"].square().mean()).backward()
            (out_m.F.square().mean() + bev_m["block8"].square().mean()).backward()
        finally:
            torch.set_num_threads(torch.get_num_threads())
        gr, gm = dict(r.named_parameters()), dict(m.named_parameters())
        worst = max(float((gr[k].grad - gm[k].grad).abs().max() / gr[k].grad.abs().max().clamp_min(1e-30)) for k in gr)
        assert worst <= 1e-5, worst  # same graph; the BEV scatter's backward sums in a different order


def test_reference_losses_equal_the_device_resident_losses():
    from oracle import me_cpu
    from lidog_b200.lidog import losses
    with rh.reference(me_cpu):
        ref = rh.load_file("utils/losses/losses.py", "losses")
        g = torch.Generator().manual_seed(0)
        for C, kitti in ((7, False), (19, True)):
            logits = torch.randn(5000, C, generator=g, dtype=torch.float64) * 2
            target = torch.randint(-1, C, (5000,), generator=g)
            target[:50] = 1
            target[50:90] = min(6, C - 1)
            for ours, theirs in (
                    (losses.dice_loss(logits, target, -1), ref.DICELoss(ignore_label=-1)(logits, target)),
                    (losses.soft_dice_loss(logits, target, -1, is_kitti=kitti),
                     ref.SoftDICELoss(ignore_label=-1, is_kitti=kitti)(logits, target))):
                # float64 logits expose the formula; the reference builds its soft targets in float32
                # (`torch.empty(t_vector.shape)`, losses.py:105), worth ~1e-8 -- anything structural would be >= 1e-3
                assert abs(float(ours) - float(theirs)) <= 1e-7, (C, float(ours), float(theirs))
            # gradients too (the training signal), in float32 as trained
            l32 = logits.float().requires_grad_(True)
            l32b = logits.float().requires_grad_(True)
            losses.soft_dice_loss(l32, target, -1, is_kitti=kitti).backward()
            ref.SoftDICELoss(ignore_label=-1, is_kitti=kitti)(l32b, target).backward()
            assert float((l32.grad - l32b.grad).abs().max()) <= 1e-6 * float(l32b.grad.abs().max())
            # a class absent from the batch and an all-ignored batch
            t2 = target.clone()
            t2[t2 == 2] = 3
            assert abs(float(losses.soft_dice_loss(logits, t2, -1)) - float(ref.SoftDICELoss(ignore_label=-1)(logits, t2))) <= 1e-7


def _reference_class(relpath, name):
    """One class of a reference file, compiled from the file where it lies (the module's top imports pull open3d /
    torchvision / the dataset tree; the class itself is self-contained numpy)."""
    import ast
    path = rh.REF + "/" + relpath
    tree = ast.parse(open(path).read(), path)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name)
    class _Np:  # the reference predates numpy 1.24: `np.int` (semantickitti_bev.py:384) is the builtin int
        int = int

        def __getattr__(self, k):
            return getattr(np, k)

    ns = {"np": _Np(), "torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


@pytest.mark.parametrize("relpath", ["utils/datasets/semantickitti_bev.py", "utils/datasets/synth4d_bev.py"])
def test_reference_bev_label_image_equals_the_device_function(relpath):
    """PC2ImgConverter.getBEVImageNew (semantickitti_bev.py:433-464 == synth4d_bev.py:478-509) vs step.bev_label_image,
    with the dataset's own argument preparation (semantickitti_bev.py:140-153, :244-249)."""
    from oracle import voxel as ov
    from lidog_b200.lidog import step
    Conv = _reference_class(relpath, "PC2ImgConverter")
    for seed, shape, bound, img in ((3, "nuscenes", 30.0, 100), (5, "kitti", 50.0, 167), (7, "nuscenes", 8.0, 20)):
        pts, lab = _scan(seed, r=60.0 if shape == "kitti" else 40.0, shape=shape)
        q, _, colab, vidx, _ = ov.sparse_quantize(pts, np.ones((len(pts), 1), np.float32), lab, -1, True, True, False, 0.05)
        bev_points = (q * 0.05).astype(np.float32)                      # semantickitti_bev.py:244
        bounds = [[-bound, bound], [-bound, bound], [-10, 8]]             # :137
        conv = Conv(imgChannel=1, xRange=bounds[0], yRange=bounds[1], zRange=bounds[2],
                    xGridSize=(bounds[0][1] - bounds[0][0]) / img, yGridSize=(bounds[1][1] - bounds[1][0]) / img,
                    zGridSize=0.3)                                        # :143-153
        ref_img, _ = conv.getBEVImageNew(bev_points, colab)
        got = step.bev_label_image(torch.from_numpy(ov.batched_coordinates([q])), torch.from_numpy(colab), 1, bound, img)
        assert ref_img.shape == (img, img)
        assert np.array_equal(ref_img.astype(np.int64), got[0].numpy()), (seed, bound)
        assert (ref_img >= 0).sum() > 50


def _reference_items(scans, me, Conv, bound, img, voxel=0.05):
    """The per-scan dictionaries the reference dataset's __getitem__ builds (semantickitti_bev.py:209-290), from the
    same calls: the shim's sparse_quantize, the reference's own PC2ImgConverter."""
    bounds = [[-bound, bound], [-bound, bound], [-10, 8]]
    conv = Conv(imgChannel=1, xRange=bounds[0], yRange=bounds[1], zRange=bounds[2], xGridSize=2 * bound / img,
                yGridSize=2 * bound / img, zGridSize=0.3)
    items = []
    for i, (pts, lab) in enumerate(scans):
        colors = np.ones((len(pts), 1), np.float32)                                         # :194
        q, _, qlab, vidx, inv = me.utils.sparse_quantize(pts, colors, labels=lab, ignore_label=-1,
                                                         quantization_size=voxel, return_index=True,
                                                         return_inverse=True)               # :232-238
        bev_points = (np.asarray(q) * voxel).astype(np.float32)                              # :244
        bl, bs = conv.getBEVImageNew(bev_points, np.asarray(qlab))                           # :249
        items.append({"coordinates": torch.as_tensor(np.asarray(q)), "features": torch.from_numpy(colors[np.asarray(vidx)]),
                      "sem_labels": torch.from_numpy(lab[np.asarray(vidx)]), "xyz": torch.from_numpy(pts[np.asarray(vidx)]),
                      "idx": torch.tensor(i), "sampled_idx": torch.as_tensor(np.asarray(vidx)),
                      "bev_labels": {"block8": torch.from_numpy(bl).long()},
                      "bev_selected_idx": {"block8": torch.from_numpy(bs).long()}})
    return items


def test_reference_collation_and_training_step_run_unchanged_and_equal_the_mirror():
    """collation.py:274-325 + trainer_lighting_2d.py:141-293,349-360 with the LiDOG configuration
    (configs/lidog/single/semantickitti.yaml: SoftDICELoss + DICELoss, weights 0.5/0.5, Adam 1e-3, warmup 0,
    clear_cache_int 1) against LidogTrainer: same batch products, same losses, same gradients, same Adam step."""
    from oracle import me_cpu
    from oracle.me_cpu.bevfn import sparse2super as o_s2s
    from lidog_b200.lidog import model as M, step
    bound, img = 8.0, 27  # 2 * 8 m / 0.05 = 320 px -> MaxPool2d(5,3,1) 106 -> two stride-2 convs -> 27 (kitti: 2000 -> 167)
    scans = [_scan(21, bound), _scan(22, bound)]
    Conv = _reference_class("utils/datasets/semantickitti_bev.py", "PC2ImgConverter")
    with rh.reference(me_cpu):
        import utils.collation.collation as coll
        import utils.models.minkunet_bev as ref
        tr2d = rh.load_file("utils/pipelines/trainer_lighting_2d.py", "trainer_lighting_2d")
        batch = coll.CollateFNSingleSourceBEVMultiLevel(device=None)(_reference_items(scans, me_cpu, Conv, bound, img))
        assert batch["source_coordinates0"].dtype == torch.float32 and batch["source_coordinates0"].shape[1] == 4
        torch.manual_seed(0)
        net_r = ref.MinkUNet34BEV(1, 7, 3, mapping_bound_2d=bound)
        state = {k: v.clone() for k, v in net_r.state_dict().items()}
        ds = rh.FakeDataset(7)
        plt = tr2d.PLTTrainer2D(model=net_r, training_dataset=ds, validation_dataset=ds, optimizer_name="Adam",
                                sem_criterion="SoftDICELoss", sem_bev_criterion="DICELoss", aux_criterion=None,
                                warmup_epochs=0, lr=1e-3, batch_size=2, weight_decay=1e-4, num_classes=7, clear_cache_int=1,
                                source_weights=[0.5, 0.5], source_domains_name=["SemanticKITTI-BEV"],
                                target_domains_name=None)
        opt = plt.configure_optimizers()
        plt.trainer.optimizers = [opt]
        torch.set_num_threads(1)  # deterministic index_put_ in the reference's sparse2super
        opt.zero_grad()
        total_r = plt.training_step(batch, 0)  # the UNCHANGED step: builds ME.SparseTensor, model, CPU losses, metrics
        total_r.backward()
        grads_r = {k: p.grad.clone() for k, p in net_r.named_parameters()}
        opt.step()
        logged = dict(plt.logged)

    # the mirror on the same scans: device-style batched voxelisation + label products, same weights
    net_m = M.MinkUNet34BEV(1, 7, ME=me_cpu, bev_fn=o_s2s, mapping_bound_2d=bound)
    net_m.load_state_dict(state)
    mir = step.LidogTrainer(net_m, num_classes=7, shape="nuscenes", ME=me_cpu)
    mir.bound, mir.bev_img = bound, img
    P, Lb = [torch.from_numpy(p) for p, _ in scans], [torch.from_numpy(l) for _, l in scans]
    coords, feats, sem, bev, cm = mir.voxelize(P, Lb)
    assert torch.equal(coords.to(torch.float32), batch["source_coordinates0"])            # batched voxel coordinates
    assert torch.equal(sem, batch["source_sem_labels0"].long())                            # first-point labels
    assert torch.equal(bev, batch["source_bev_labels0"]["block8"])                         # BEV label images
    total_m, l3, l2 = mir.forward_loss(coords, feats, sem, bev, len(P), cm)
    mir.optimizer.zero_grad(set_to_none=True)
    total_m.backward()
    assert abs(float(total_m) - float(total_r)) <= 1e-6, (float(total_m), float(total_r))
    key3d = "training/SemanticKITTI-BEV/sem_loss0"
    assert abs(logged[key3d] - float(l3)) <= 1e-6 and abs(logged["training/SemanticKITTI-BEV/bev_loss0"] - float(l2)) <= 1e-6
    worst = 0.0
    for k, p in net_m.named_parameters():
        g, gr = p.grad, grads_r[k]
        if float(gr.norm()) > 1e-10:
            worst = max(worst, float((g - gr).norm() / gr.norm()))
    assert worst <= 1e-4, worst  # same arithmetic; masked-sum losses instead of boolean indexing reorder float sums
    mir.optimizer.step()
    moved = max(float((p.detach() - net_r.state_dict()[k]).abs().max()) for k, p in net_m.named_parameters())
    assert moved <= 5e-4, moved  # both took the same Adam(lr 1e-3, wd 1e-4) step (|update| ~ 1e-3 each)


def _voxelised_sample(seed, r, idx):
    from oracle import voxel as ov
    pts, lab = _scan(seed, r)
    q, f, _, vidx, inv = ov.sparse_quantize(pts, np.ones((len(pts), 1), np.float32), lab, -1, True, True, False, 0.05)
    return dict(coordinates=torch.from_numpy(q), xyz=torch.from_numpy(pts[vidx]), features=torch.from_numpy(f),
                sem_labels=torch.from_numpy(lab[vidx]), sampled_idx=torch.from_numpy(vidx), idx=torch.tensor(idx),
                inverse_map=torch.from_numpy(inv))


def test_reference_mix3d_merge_equals_the_device_merge():
    """`merge_data` (utils/datasets/mix3D.py:43-87), run unchanged on the shim, vs datapath.mix3d_merge."""
    import ast
    import types
    from oracle import me_cpu
    from lidog_b200.lidog import datapath
    src = open(rh.REF + "/utils/datasets/mix3D.py").read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "merge_data")
    ns = {"np": np, "torch": torch, "ME": me_cpu}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), rh.REF + "/utils/datasets/mix3D.py", "exec"), ns)
    s0, s1 = _voxelised_sample(31, 9.0, 0), _voxelised_sample(32, 9.0, 1)
    ref = ns["merge_data"](types.SimpleNamespace(voxel_size=0.05, ignore_label=-1), s0, s1)
    got = datapath.mix3d_merge(s0, s1, 0.05, -1, ME=me_cpu)
    assert len(ref["coordinates"]) < len(s0["coordinates"]) + len(s1["coordinates"])  # overlapping voxels merged
    for k in ("coordinates", "features", "sem_labels", "sampled_idx", "xyz", "idx"):
        assert torch.equal(torch.as_tensor(ref[k]), torch.as_tensor(got[k])), k
    # the float32 round trip is in there: the merged set is not the plain union of the two integer sets
    union = np.unique(np.concatenate([s0["coordinates"].numpy(), s1["coordinates"].numpy()]), axis=0)
    assert not np.array_equal(np.unique(np.asarray(ref["coordinates"]), axis=0), union)

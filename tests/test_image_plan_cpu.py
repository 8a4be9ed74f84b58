"""Deferred augmentation: the ops of textboost_b200.augment run on an ImagePlan record primitives
instead of calling PIL; executing the record with the pinned oracles (tests/plan_standin.py) must reproduce, byte for
byte, what the same ops — same seeds, same random streams — produce eagerly with PIL / torchvision, and the captions
and the state of the random streams must be identical.  The GPU executor runs the same primitives."""
import random

import numpy as np
import pytest
import torch
from PIL import Image

import make_augment_golden as G
import plan_standin


def _plan_of(img):
    from textboost_b200.image_plan import ImagePlan
    return ImagePlan(torch.from_numpy(np.array(img, dtype=np.uint8)))


DEFERRED_OPS = ["adjust_scale", "horizontal_flip", "horizontal_translate", "grayscale", "crop", "square_photo_collage"]


@pytest.mark.parametrize("name", DEFERRED_OPS)
@pytest.mark.parametrize("size", [(64, 64), (96, 72), (50, 81)])
def test_each_op_recorded_equals_eager(name, size):
    from textboost_b200 import augment
    fn = getattr(augment, name)
    for inversion in (False, True):
        for seed in range(5):
            img = G.make_image(size, seed)
            G.seed_all(seed * 7 + 1)
            eager, prompt_e = fn(img, "a photo of a <sks> dog", inversion)
            after_e = (float(np.random.random()), random.random())
            G.seed_all(seed * 7 + 1)
            plan, prompt_p = fn(_plan_of(img), "a photo of a <sks> dog", inversion)
            after_p = (float(np.random.random()), random.random())
            assert prompt_e == prompt_p and after_e == after_p
            assert plan.size == eager.size
            assert np.array_equal(plan_standin.run(plan), np.asarray(eager)), (name, size, inversion, seed, plan)


def test_ops_outside_the_pipeline_refuse_plans():
    from textboost_b200 import augment
    plan = _plan_of(G.make_image((32, 32), 0))
    for name in ("rotate", "adjust_brightness", "jpeg_compression"):
        with pytest.raises(NotImplementedError):
            getattr(augment, name)(plan, "a dog", False)
    with pytest.raises(NotImplementedError):
        plan.transpose(Image.FLIP_TOP_BOTTOM)
    with pytest.raises(NotImplementedError):
        plan.crop((-1, 0, 10, 10))
    assert plan.resize((32, 32), Image.BICUBIC) is plan and plan.copy() is plan
    assert plan.pad_edge(2, 3).size == (36, 38) and plan.collage(3).size == (96, 96)


@pytest.mark.parametrize("cfg", range(len(G.PIPES)))
def test_pipeline_streams_recorded_equal_eager(cfg):
    """One continuous 24-image stream per PairedAugmentation configuration (the golden configurations): captions,
    bytes and final random-stream state agree between eager PIL execution and record-then-execute."""
    from textboost_b200 import augment
    kw = G.PIPES[cfg]
    eager_pipe, plan_pipe = augment.PairedAugmentation(**kw), augment.PairedAugmentation(**kw)
    G.seed_all(123)
    eager = [eager_pipe(G.make_image(G.SIZES[0], i), "a <sks> dog") for i in range(24)]
    end_e = (float(np.random.random()), random.random())
    G.seed_all(123)
    plans = [plan_pipe(_plan_of(G.make_image(G.SIZES[0], i)), "a <sks> dog") for i in range(24)]
    end_p = (float(np.random.random()), random.random())
    assert end_e == end_p
    n_ops = 0
    for (img, prompt, mask), (plan, prompt_p, mask_p) in zip(eager, plans):
        assert prompt == prompt_p and mask is None and mask_p is None
        assert np.array_equal(plan_standin.run(plan), np.asarray(img)), plan
        n_ops += len(plan.ops)
    assert n_ops > 0 or kw["p"] == 0

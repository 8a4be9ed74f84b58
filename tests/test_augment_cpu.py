"""Image half of the data pipeline against golden vectors generated from the reference itself
(tests/golden/make_augment_golden.py imports /root/reference/textboost/{augment/paired_augmentation.py,dataset.py}):
same seeds -> the same prompts verbatim, the same image bytes (SHA-256), and the same state of the random streams
afterwards (one more draw from each is compared)."""
import json
import os

import pytest

import make_augment_golden as G

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def _same_libs(meta):
    import numpy
    import PIL
    import torchvision
    return (meta["pillow"], meta["torchvision"], meta["numpy"]) == (PIL.__version__, torchvision.__version__,
                                                                     numpy.__version__)


def test_every_augmentation_op_matches_reference_golden():
    from textboost_b200 import augment
    gold = _load("augment_golden.json")
    if not _same_libs(gold["meta"]):
        pytest.skip("Pillow / torchvision / numpy differ from the versions the golden bytes were made with")
    ours = G.run_ops(augment)
    assert len(ours) == len(gold["ops"]) == len(G.OPS) * 2 * 2 * 4
    for o, g in zip(ours, gold["ops"]):
        assert o == g, (g["op"], g["size"], g["inversion"], g["seed"], o["prompt"], g["prompt"])


def test_paired_augmentation_streams_match_reference_golden():
    from textboost_b200 import augment
    gold = _load("augment_golden.json")
    if not _same_libs(gold["meta"]):
        pytest.skip("library versions differ from the golden's")
    ours = G.run_pipes(augment)
    assert len(ours) == len(gold["pipes"])
    for o, g in zip(ours, gold["pipes"]):
        assert o == g, (g.get("config"), g.get("i"), o.get("prompt"), g.get("prompt"))
    prompts = {g["prompt"] for g in gold["pipes"] if "prompt" in g}
    assert any("<zoom" in p or "<crop>" in p or "<left_0>" in p or "<right_0>" in p for p in prompts)
    assert any("<collage_0>" in p for p in prompts) and any("grayscale" in p for p in prompts)


def test_paired_augmentation_interface():
    from PIL import Image
    from textboost_b200.augment import PairedAugmentation, horizontal_flip
    with pytest.raises(AssertionError):
        PairedAugmentation(hflip="maybe")
    pipe = PairedAugmentation(hflip="inversion", inversion=True, ops="style")
    assert pipe.geometric_ops == [horizontal_flip] and pipe.other_ops == [] and len(pipe.color_ops) == 1
    assert PairedAugmentation(hflip="TRUE").hflip is True
    with pytest.raises(AssertionError):
        pipe("not an image", "a dog")
    img = Image.new("RGB", (32, 32), (10, 200, 30))
    out, prompt, mask = PairedAugmentation(p=0.0, color_prob=0.0)(img, "a dog")
    assert out is img and prompt == "a dog" and mask is None


def test_textboost_dataset_matches_reference_golden():
    from textboost_b200 import augment, dataset
    from textboost_b200.synthetic import LiteralTokenizer
    gold = _load("dataset_golden.json")
    if not _same_libs(gold["meta"]):
        pytest.skip("library versions differ from the golden's")
    ours = G.run_dataset(dataset, augment, LiteralTokenizer)
    assert len(ours) == len(gold["datasets"]) == 5
    for o, g in zip(ours, gold["datasets"]):
        assert o["len"] == g["len"]
        for i, (oi, gi) in enumerate(zip(o["items"], g["items"])):
            assert oi == gi, (g["config"], i, oi, gi)
        assert o["batch"] == g["batch"] and o["next_draws"] == g["next_draws"], g["config"]
    assert "class_image" in gold["datasets"][3]["items"][0]["keys"]


def test_dataset_errors_and_reference_import_names(tmp_path):
    from textboost_b200 import dataset, prompts
    assert dataset.Wrapper is prompts.ShardedStream and dataset.PriorDataset is prompts.PriorPrompts
    assert dataset.InstructPix2PixDataset is prompts.HumanPromptSource
    with pytest.raises(ValueError):
        dataset.get_images_path(tmp_path / "missing")
    for n in ("b.png", "a.png", "c.png"):
        (tmp_path / n).write_bytes(b"")
    assert [p.name for p in dataset.get_images_path(tmp_path, 2)] == ["a.png", "b.png"]
    assert len(dataset.get_images_path(tmp_path)) == 3


def test_cli_image_batches_shapes_sharding_and_determinism(tmp_path):
    """train_textboost.build_image_batches = the reference's train dataloader wiring (train_textboost.py:856-890)."""
    import random
    import numpy as np
    import torch
    import train_textboost as T
    from textboost_b200.synthetic import LiteralTokenizer
    d = tmp_path / "imgs"
    d.mkdir()
    for i, size in enumerate([(80, 64), (64, 64), (70, 90), (66, 66), (64, 100)]):
        G.make_image(size, i).save(d / f"{i:02d}.png")

    def batches(rank, world, n, extra=()):
        args = T.parse_args(["--pretrained_model_name_or_path", "x", "--instance_data_dir", str(d), "--resolution", "32",
                             "--train_batch_size", "2", "--augment", "pda", "--augment_inversion", "--seed", "5",
                             "--template", "textboost", *extra])
        args.concepts_list = [{"instance_token": ["<dog>"], "instance_data_dir": str(d)}]
        G.seed_all(5)
        it = T.build_image_batches(args, LiteralTokenizer(), rank, world)
        return [next(it) for _ in range(n)]

    a = batches(0, 1, 4)
    assert T.RUN_INFO["instance_images"] == 5
    for b in a:
        assert b["pixel_values"].shape == (2, 3, 32, 32) and b["pixel_values"].dtype == torch.float32
        assert -1.0 <= b["pixel_values"].min() and b["pixel_values"].max() <= 1.0
        assert b["input_ids"].shape == (2, 77) and len(b["attention_mask"]) == 2
    again = batches(0, 1, 4)
    assert all(torch.equal(x["pixel_values"], y["pixel_values"]) and torch.equal(x["input_ids"], y["input_ids"])
               for x, y in zip(a, again))
    # the prompt carries the list-valued instance token verbatim (SURVEY.md trap 14) through LiteralTokenizer ids
    tok = LiteralTokenizer()
    assert tok._word("['<dog>']") in a[0]["input_ids"][0].tolist() or any(
        tok._word("['<dog>']") in row for b in a for row in b["input_ids"].tolist())
    # two ranks see disjoint halves of each shuffled epoch (5 images padded to 6, 3 per rank)
    order = np.arange(5)
    np.random.default_rng(seed=5).shuffle(order)
    idx = np.concatenate((order, order[:1]))
    assert len(set(idx[0::2][:3]) | set(idx[1::2][:3])) == 5
    r0, r1 = batches(0, 2, 2, ("--augment", "none", "--center_crop")), batches(1, 2, 2, ("--augment", "none",
                                                                                           "--center_crop"))
    assert not torch.equal(r0[0]["pixel_values"], r1[0]["pixel_values"])
    with __import__("pytest").raises(NotImplementedError):
        batches(0, 1, 1, ("--augment", "custom_diff"))
    random.seed(0)


def test_cli_image_batches_with_image_prior(tmp_path):
    """--with_image_prior: class examples are appended after the instance ones ([2B, ...]), prompts from the class
    token through the same template draw (dataset.py:393-418, 420-459)."""
    import train_textboost as T
    from textboost_b200.synthetic import LiteralTokenizer
    inst, cls = tmp_path / "dog", tmp_path / "class"
    inst.mkdir()
    cls.mkdir()
    for i in range(2):
        G.make_image((70 + 10 * i, 64), i).save(inst / f"{i}.png")
    for i in range(3):
        G.make_image((64, 72), 10 + i).save(cls / f"{i}-a_photo_of_dog.png")
    args = T.parse_args(["--pretrained_model_name_or_path", "x", "--instance_data_dir", str(inst), "--resolution", "32",
                         "--train_batch_size", "2", "--with_image_prior", "--class_data_dir", str(cls),
                         "--class_token", "dog", "--template", "a {}", "--seed", "1"])
    args.concepts_list = [{"instance_token": "<dog>", "instance_data_dir": str(inst)}]
    G.seed_all(1)
    b = next(T.build_image_batches(args, LiteralTokenizer(), 0, 1))
    assert b["pixel_values"].shape == (4, 3, 32, 32) and b["input_ids"].shape == (4, 77)
    assert len(b["attention_mask"]) == 4
    tok = LiteralTokenizer()
    assert b["input_ids"][0].tolist() == tok("a <dog>").input_ids[0].tolist()
    # --class_token is nargs="+" in the reference (train_textboost.py:96-101), so the class prompt formats a LIST
    assert b["input_ids"][2].tolist() == b["input_ids"][3].tolist() == tok("a ['dog']").input_ids[0].tolist()
    with __import__("pytest").raises(ValueError):
        T._unsupported(T.parse_args(["--pretrained_model_name_or_path", "x", "--with_image_prior", "--class_data_dir",
                                     str(cls), "--class_token", "dog", "--synthetic_data"]))


def test_image_stream_under_real_dataloader_workers(tmp_path):
    """Wrapper / ShardedStream as an IterableDataset inside torch's DataLoader with worker processes: every
    (rank, worker) pair takes its own stride of the shuffled epoch (dataset.py:846-870), batches keep their shape."""
    import torch
    import train_textboost as T
    from textboost_b200 import dataset, prompts
    from textboost_b200.synthetic import LiteralTokenizer
    d = tmp_path / "imgs"
    d.mkdir()
    for i in range(8):
        G.make_image((40 + i, 40), i).save(d / f"{i:02d}.png")
    ds = dataset.TextBoostDataset([{"instance_token": "<dog>", "instance_data_dir": str(d)}], LiteralTokenizer(),
                                  size=32, center_crop=True, template="a {}")
    seen = {}
    for rank in (0, 1):
        stream = prompts.ShardedStream(ds, drop_last=False, rank=rank, world_size=2).shuffle(seed=3)  # one epoch
        loader = torch.utils.data.DataLoader(stream, batch_size=1, num_workers=2,
                                             collate_fn=lambda ex: dataset.TextBoostDataset.collate_fn(ex, False))
        seen[rank] = [b["pixel_values"] for b in loader]
        assert all(p.shape == (1, 3, 32, 32) for p in seen[rank])
    # 8 images over 2 ranks x 2 workers: each rank sees 4 distinct images, the two ranks together all 8
    def key(p):
        return round(float(p.double().sum()), 4)
    k0, k1 = {key(p) for p in seen[0]}, {key(p) for p in seen[1]}
    assert len(seen[0]) == len(seen[1]) == 4 and len(k0) == len(k1) == 4 and not (k0 & k1)
    args = T.parse_args(["--pretrained_model_name_or_path", "x", "--instance_data_dir", str(d), "--resolution", "32",
                         "--train_batch_size", "3", "--dataloader_num_workers", "2", "--center_crop"])
    args.concepts_list = [{"instance_token": "<dog>", "instance_data_dir": str(d)}]
    it = T.build_image_batches(args, LiteralTokenizer(), 0, 1)
    assert [next(it)["pixel_values"].shape for _ in range(4)] == [(3, 3, 32, 32)] * 4  # endless: repeat() wraps epochs
    del it

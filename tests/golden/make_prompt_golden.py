"""Generates tests/golden/prompts_golden.json by running the REFERENCE's own prompt pipeline classes
(/root/reference/textboost/dataset.py: Wrapper, InstructPix2PixDataset, PriorDataset, the template tables) in this
container.  /root/reference does not exist on the GPU box, so only the outputs travel.

    python tests/golden/make_prompt_golden.py
"""
import hashlib
import itertools
import json
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

JSONL = [{"input": f"a photo of thing number {i}", "output": None if i % 3 == 0 else ("NONE" if i % 3 == 1 else f"thing {i} at night")}
         for i in range(23)]


def write_jsonl(path):
    with open(path, "w") as f:
        for row in JSONL:
            f.write(json.dumps(row) + "\n")


def main():
    import torch.utils.data
    from textboost import dataset as R
    from textboost_b200.synthetic import LiteralTokenizer

    out = {"jsonl": JSONL, "reference": "textboost/dataset.py:13-93, 161-269, 827-882"}
    out["template_sha"] = {k: hashlib.sha256("\n".join(v).encode()).hexdigest() for k, v in
                           (("imagenet_small", R.imagenet_templates_small),
                            ("imagenet_style_small", R.imagenet_style_templates_small),
                            ("textboost", R.textboost_templates))}
    tok = LiteralTokenizer()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "p.jsonl")
        write_jsonl(path)
        src = R.InstructPix2PixDataset(tok, path)
        out["source_data"] = list(src.data)
        src5 = R.InstructPix2PixDataset(tok, path, num_samples=5)
        out["source_data_5"] = list(src5.data)
        prior = R.PriorDataset(src, tok, additional_template="imagenet_small", additional_category=["dog", "cat"],
                               null_prob=0.1)
        out["prior_template_data"] = prior.template_data
        random.seed(1234)
        out["prior_draws"] = [prior[i % len(prior)]["prompt"] for i in range(300)]
        prior2 = R.PriorDataset(src, tok, additional_template="a {} toy", additional_category="robot", null_prob=0.3,
                                template_prob=0.2)
        random.seed(7)
        out["prior_draws_2"] = [prior2[(3 * i) % len(prior2)]["prompt"] for i in range(200)]

        # Wrapper index sequences (rank x worker sharding is emulated through torch.distributed / worker_info mocks)
        class Idx:  # source[i] -> i
            def __init__(self, n):
                self.n = n

            def __len__(self):
                return self.n

            def __getitem__(self, i):
                return int(i)

        cases = []
        for (n, world, rank, workers, wid, shuffle, seed, drop_last, take) in [
                (10, 1, 0, 0, 0, False, None, True, 25), (10, 1, 0, 0, 0, True, 42, True, 35),
                (23, 2, 1, 0, 0, True, 42, True, 40), (23, 2, 0, 0, 0, True, 42, False, 40),
                (23, 4, 3, 2, 1, True, 5, True, 30), (7, 8, 2, 0, 0, True, 0, False, 12), (16, 8, 5, 0, 0, True, 9, True, 10)]:
            w = R.Wrapper(Idx(n), drop_last=drop_last)
            if shuffle:
                w = w.shuffle(seed=seed)
            w = w.repeat()
            import torch.distributed as dist
            orig = (dist.is_initialized, dist.get_world_size, dist.get_rank, torch.utils.data.get_worker_info)
            dist.is_initialized = lambda: world > 1
            dist.get_world_size = lambda *a, **k: world
            dist.get_rank = lambda *a, **k: rank

            class WI:
                num_workers = workers
                id = wid
            torch.utils.data.get_worker_info = (lambda: WI) if workers else (lambda: None)
            try:
                seq = list(itertools.islice(iter(w), take))
            finally:
                dist.is_initialized, dist.get_world_size, dist.get_rank, torch.utils.data.get_worker_info = orig
            cases.append({"n": n, "world": world, "rank": rank, "workers": workers, "worker_id": wid,
                          "shuffle": shuffle, "seed": seed, "drop_last": drop_last, "indices": seq})
        out["wrapper_cases"] = cases
    with open(os.path.join(HERE, "prompts_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote prompts_golden.json:", len(out["prior_draws"]), "draws,", len(cases), "wrapper cases")


if __name__ == "__main__":
    main()

"""CPU parity tests of the prompt side of the data pipeline (textboost_b200/prompts.py) against golden outputs of
the reference's own classes (tests/golden/prompts_golden.json, written by tests/golden/make_prompt_golden.py from
/root/reference/textboost/dataset.py): template tables, the JSONL prompt source, the null / template / human prior
sampler with the same `random` draws, and the rank x worker sharded, shuffled, repeated index stream."""
import hashlib
import itertools
import json
import os
import random

import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "prompts_golden.json")


def _gold():
    with open(GOLDEN) as f:
        return json.load(f)


def _write_jsonl(path, rows):
    with open(path, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


def test_templates_match_reference_tables():
    from textboost_b200 import prompts as P
    g = _gold()
    for k, sha in g["template_sha"].items():
        assert hashlib.sha256("\n".join(P.TEMPLATES[k]).encode()).hexdigest() == sha
    assert P.resolve_template("photo of {} on a beach") == ["photo of {} on a beach"]
    random.seed(3)
    p = P.instance_prompt(P.TEMPLATES["textboost"], "<dog> dog")
    assert p.endswith("<dog> dog") and p.replace("<dog> dog", "{}") in P.TEMPLATES["textboost"]
    # the reference formats the template with the LIST of placeholder strings (SURVEY.md trap 14)
    assert P.instance_prompt(["a {}"], ["<dog>"]) == "a ['<dog>']"


def test_prompt_source_and_prior_sampler_match_reference(tmp_path):
    from textboost_b200 import prompts as P
    from textboost_b200.synthetic import LiteralTokenizer
    g = _gold()
    path = str(tmp_path / "p.jsonl")
    _write_jsonl(path, g["jsonl"])
    tok = LiteralTokenizer()
    src = P.HumanPromptSource(tok, path)
    assert src.data == g["source_data"]
    assert P.HumanPromptSource(tok, path, num_samples=5).data == g["source_data_5"]
    item = src[2]
    assert item["prompt"] == src.data[2] and item["input_ids"].shape == (1, 77) and item["input_ids"][0, 0] == 49406
    prior = P.PriorPrompts(src, tok, additional_template="imagenet_small", additional_category=["dog", "cat"], null_prob=0.1)
    assert prior.template_data == g["prior_template_data"]
    random.seed(1234)
    assert [prior[i % len(prior)]["prompt"] for i in range(300)] == g["prior_draws"]
    prior2 = P.PriorPrompts(src, tok, additional_template="a {} toy", additional_category="robot", null_prob=0.3,
                            template_prob=0.2)
    random.seed(7)
    assert [prior2[(3 * i) % len(prior2)]["prompt"] for i in range(200)] == g["prior_draws_2"]
    frac_null = sum(p == "" for p in g["prior_draws"]) / 300
    assert 0.04 < frac_null < 0.18  # null_prob = 0.1
    batch = P.PriorPrompts.collate_fn([prior[0], prior[1], prior[2]])
    assert batch["input_ids"].shape == (3, 77) and len(batch["prompt"]) == 3
    # an empty prompt tokenises to BOS, EOS, EOS...: the row TextBoostModel.forward overwrites (text_encoder.py:71)
    ids = P.tokenize_prompt(tok, "").input_ids
    assert ids[0, 0] == 49406 and ids[0, 1] == 49407


def test_sharded_stream_matches_reference_wrapper():
    from textboost_b200 import prompts as P

    class Idx:
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __getitem__(self, i):
            return int(i)

    class WI:
        pass

    for c in _gold()["wrapper_cases"]:
        s = P.ShardedStream(Idx(c["n"]), drop_last=c["drop_last"], rank=c["rank"], world_size=c["world"])
        if c["shuffle"]:
            s = s.shuffle(seed=c["seed"])
        s = s.repeat()
        orig = torch.utils.data.get_worker_info
        if c["workers"]:
            wi = WI()
            wi.num_workers, wi.id = c["workers"], c["worker_id"]
            torch.utils.data.get_worker_info = lambda: wi
        try:
            got = list(itertools.islice(iter(s), len(c["indices"])))
        finally:
            torch.utils.data.get_worker_info = orig
        assert got == c["indices"], c
    # the shards of one epoch partition the (trimmed) index set
    n, world = 23, 4
    seen = []
    for r in range(world):
        s = P.ShardedStream(Idx(n), rank=r, world_size=world).shuffle(seed=1)
        seen += list(s.indices())
    assert len(seen) == n - n % world and len(set(seen)) == len(seen)

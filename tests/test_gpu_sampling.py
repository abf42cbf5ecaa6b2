"""GPU: the reference's generation config (inference/inference.py:41-42, 426-451: do_sample=True, temperature=0.9,
repetition_penalty=1.05; processors + multinomial at generate/streaming_generate_qwen.py:73-99) inside the lm_head kernel
(`svlm_lm_head_sample`), through the C ABI.

* processed scores (HF `output_scores`): bit-exact against HF's own `RepetitionPenaltyLogitsProcessor` and
  `TemperatureLogitsWarper` applied to the kernel's raw bf16 logits; raw logits against the oracle's lm_head within tolerance;
* Gumbel-max sampling: the empirical distribution of 6000 draws equals softmax(processed scores) within 4.5 sigma per token,
  draws are a pure function of (seed, step), the device-side step counter advances by itself;
* the sampled token joins the repetition-penalty set on the device.
"""
import math

import pytest
import torch

from oracle.vit import rmsnorm as o_rmsnorm

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def abi():
    from streaming_vlm_b200 import _abi
    _abi.build()
    return _abi


def _setup(vocab=4096, hidden=512, seed=0):
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(vocab, hidden, generator=g) * 0.05).to(torch.bfloat16)
    x = torch.randn(hidden, generator=g).to(torch.bfloat16)
    nw = (1.0 + 0.1 * torch.randn(hidden, generator=g)).to(torch.bfloat16)
    return w, x, nw


def _state():
    z = lambda n, dt: torch.zeros(n, dtype=dt, device=DEV)
    return dict(key=z(1, torch.int64), tok=z(1, torch.int64), cnt=z(3, torch.int32), rng=z(4, torch.int32))


def _run(abi, w, x, nw, st, suppress=None, seen=None, pen=1.0, temp=1.0, sample=False, scores=None):
    abi.lm_head_sample(w, x, nw, 1e-6, suppress, st["key"], st["tok"], st["cnt"][0:1], st["cnt"][1:2], st["cnt"][2:3], seen=seen,
                       repetition_penalty=pen, temperature=temp, do_sample=sample, rng_state=st["rng"], scores_out=scores)


def test_processed_scores_equal_hf_processors(abi):
    from transformers.generation.logits_process import RepetitionPenaltyLogitsProcessor, TemperatureLogitsWarper
    w, x, nw = _setup()
    vocab = w.shape[0]
    wd, xd, nwd = w.to(DEV), x.to(DEV), nw.to(DEV)
    st = _state()
    raw = torch.empty(vocab, dtype=torch.float32, device=DEV)
    _run(abi, wd, xd, nwd, st, scores=raw)
    # raw logits: bf16 lm_head output of the normalised hidden state (model_forward.py:316-317), upcast (streaming_generate_qwen.py:73)
    h = o_rmsnorm(x[None], nw, 1e-6)
    want_raw = torch.nn.functional.linear(h, w)[0].float()
    assert (raw.cpu() - want_raw).abs().max() < 2e-2 * want_raw.abs().max()
    assert int(st["tok"]) == int(torch.argmax(raw))          # greedy: first index of the maximum
    # processors on the kernel's own raw logits: bit-exact
    ids = torch.tensor([[5, 17, 17, 900, 4095, 33, int(torch.argmax(raw)), int(torch.argmin(raw))]])
    seen = torch.zeros(vocab, dtype=torch.uint8, device=DEV)
    abi.mark_seen(ids[0].to(DEV), ids.shape[1], seen)
    assert seen.sum().item() == len(set(ids[0].tolist()))
    suppress = torch.zeros(vocab, dtype=torch.uint8, device=DEV)
    suppress[100:200] = 1
    got = torch.empty(vocab, dtype=torch.float32, device=DEV)
    _run(abi, wd, xd, nwd, st, suppress=suppress, seen=seen, pen=1.05, temp=0.9, scores=got)
    # HF's processors ON THE GPU, where the reference evaluates them (`tensor / python_scalar` is a reciprocal multiply there)
    want = TemperatureLogitsWarper(0.9)(ids.to(DEV), RepetitionPenaltyLogitsProcessor(1.05)(ids.to(DEV), raw[None].clone()))[0].cpu()
    want[100:200] = float("-inf")
    assert torch.equal(got.cpu().view(torch.int32), want.view(torch.int32)), "processed scores differ from HF's processors"
    on_cpu = TemperatureLogitsWarper(0.9)(ids, RepetitionPenaltyLogitsProcessor(1.05)(ids, raw.cpu()[None].clone()))[0]
    on_cpu[100:200] = float("-inf")
    keep = torch.isfinite(on_cpu)
    assert (got.cpu()[keep] - on_cpu[keep]).abs().max() <= 2 * torch.finfo(torch.float32).eps * on_cpu[keep].abs().max()   # true division: <= 1 ulp apart
    assert int(st["tok"]) == int(torch.argmax(want))
    assert st["cnt"].tolist() == [2, 3, 2]                   # two launches advanced kv_len / next_pos twice; kv_after = kv_len + 1
    assert int(seen[int(st["tok"])]) == 1                    # the selected token joined the penalty set on the device


def test_gumbel_max_sampling_follows_softmax(abi):
    w, x, nw = _setup(seed=1)
    vocab = w.shape[0]
    wd, xd, nwd = w.to(DEV), x.to(DEV), nw.to(DEV)
    allowed = [3, 700, 701, 1500, 2999, 4000]
    suppress = torch.ones(vocab, dtype=torch.uint8, device=DEV)
    suppress[torch.tensor(allowed)] = 0
    st = _state()
    sc = torch.empty(vocab, dtype=torch.float32, device=DEV)
    temp = 0.5                                                # sharpen a little: the six probabilities should differ clearly
    _run(abi, wd, xd, nwd, st, suppress=suppress, temp=temp, scores=sc)
    p = torch.softmax(sc.cpu()[allowed].double(), dim=0)
    n_draws = 6000
    st["rng"].copy_(torch.tensor(abi.rng_words(0x1234567890ABCDEF), dtype=torch.int32))
    toks = torch.empty(n_draws, dtype=torch.int64, device=DEV)
    for i in range(n_draws):
        _run(abi, wd, xd, nwd, st, suppress=suppress, temp=temp, sample=True)
        toks[i] = st["tok"][0]
    assert int(st["rng"][2]) == n_draws                       # the step word advanced on the device, once per draw
    toks = toks.cpu()
    assert set(toks.tolist()) <= set(allowed)
    for a, pa in zip(allowed, p.tolist()):
        f = (toks == a).double().mean().item()
        sigma = math.sqrt(pa * (1 - pa) / n_draws)
        assert abs(f - pa) < 4.5 * sigma + 1e-4, f"token {a}: frequency {f:.4f} vs probability {pa:.4f}"
    # consecutive draws are independent: P(two successive draws differ) = 1 - sum p^2
    assert 0.02 < p.max().item() < 0.9, p
    change = (toks[1:] != toks[:-1]).double().mean().item()
    assert abs(change - (1.0 - (p * p).sum().item())) < 0.03, (change, p)
    # a draw is a pure function of (seed, step)
    st2 = _state()
    st2["rng"].copy_(torch.tensor(abi.rng_words(0x1234567890ABCDEF), dtype=torch.int32))
    for i in range(16):
        _run(abi, wd, xd, nwd, st2, suppress=suppress, temp=temp, sample=True)
        assert int(st2["tok"]) == int(toks[i])
    st2["rng"].copy_(torch.tensor(abi.rng_words(99), dtype=torch.int32))
    other = []
    for i in range(64):
        _run(abi, wd, xd, nwd, st2, suppress=suppress, temp=temp, sample=True)
        other.append(int(st2["tok"]))
    assert other != toks[:64].tolist()


def test_sampling_needs_rng_state_and_positive_temperature(abi):
    w, x, nw = _setup(vocab=256)
    wd, xd, nwd = w.to(DEV), x.to(DEV), nw.to(DEV)
    st = _state()
    with pytest.raises(RuntimeError, match="RNG state"):
        abi.lm_head_sample(wd, xd, nwd, 1e-6, None, st["key"], st["tok"], st["cnt"][0:1], st["cnt"][1:2], st["cnt"][2:3], do_sample=True)
    with pytest.raises(RuntimeError, match="must be positive"):
        _run(abi, wd, xd, nwd, st, temp=0.0)


def test_engine_sampling_lookahead_equals_plain_loop():
    """The look-ahead decode loop (next step enqueued before the host saw the token, rolled back on <|im_end|>, RNG step word
    included) produces the same SAMPLED stream as the plain read-then-step loop for the same seeds."""
    from streaming_vlm_b200 import synthetic
    from streaming_vlm_b200.config import TINY
    from streaming_vlm_b200.engine import IM_END, GenConfig, StreamingEngine
    geo = TINY
    w_cpu = synthetic.random_weights(geo, seed=0)
    suppress = torch.ones(geo.vocab, dtype=torch.bool)
    suppress[torch.tensor([IM_END, 1000, 1001, 1002, 1003, 1004])] = False
    engines = [StreamingEngine({k: v.clone() for k, v in w_cpu.items()}, geo, capacity=4096, window_size=4, text_round=4, text_sink=16,
                               text_sliding_window=24, suppress=suppress, use_graph=ug, prefill_bucket=1,
                               gen=GenConfig(do_sample=True, repetition_penalty=1.05, temperature=0.9)) for ug in (True, False)]
    early, distinct = 0, set()
    for i in range(12):
        px, grid = synthetic.patchify(synthetic.frames(i, 112, 112), geo)
        prompt = synthetic.round_prompt(i, 16, 6)
        ra, rb = (e.step(prompt, px.to("cuda"), grid) for e in engines)
        assert ra["ids"] == rb["ids"], f"chunk {i}"
        assert int(engines[0].cache.next_pos) == int(engines[1].cache.next_pos)
        early += len(ra["new_tokens"]) < engines[0].max_new_tokens
        distinct |= set(ra["new_tokens"])
    assert early >= 2 and len(distinct) >= 4, (early, distinct)
    n = engines[0].cache.seq_len
    a, b = (e.cache.arena[e.cache.cur, :, :, :, :n] for e in engines)
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))

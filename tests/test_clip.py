"""Text conditioning: the OpenCLIP ViT-H-14 text tower of FrozenOpenCLIPEmbedder
(modelscope/clip_hardcode.py:59-422).  open_clip is a third-party dependency that is not installed here and not pinned by the
reference: the oracle restates its published block structure and is pinned against torch's own nn.MultiheadAttention /
nn.LayerNorm / nn.GELU modules (CPU test below); the GPU tests compare the library path with that oracle."""
import pytest
import torch

from oracle import clip_oracle as CO, unet_oracle as UO


def test_clip_oracle_matches_torch_multihead_attention_modules():
    cfg = CO.ClipConfig(width=128, heads=2, layers=4, layers_run=3, context=77, vocab=300)
    W = UO.make_weights(CO.clip_param_specs(cfg), seed=4)
    tower = CO.TorchTextTower(cfg)
    assert set(tower.state_dict()) == set(W)
    tower.load_state_dict(W)
    tok = torch.randint(0, cfg.vocab, (3, cfg.context), generator=torch.Generator().manual_seed(1))
    ref = tower(tok)
    out = CO.clip_text_forward(W, cfg, tok)
    assert torch.allclose(out, ref, rtol=0, atol=2e-5), (out - ref).abs().max()
    # causality: changing a later token leaves every earlier position untouched
    tok2 = tok.clone()
    tok2[:, 40] = (tok2[:, 40] + 7) % cfg.vocab
    out2 = CO.clip_text_forward(W, cfg, tok2)
    assert torch.equal(out[:, :40], out2[:, :40]) and not torch.equal(out[:, 40:], out2[:, 40:])


class FakeTokenizer(object):
    """Deterministic stand-in for the BPE vocabulary (open_clip's is not available offline): one token per word."""
    encoder = {'<start_of_text>': 298, '<end_of_text>': 299, ',</w>': 5}

    def encode(self, text):
        return [5 if w == ',' else 10 + (sum(ord(c) for c in w) % 250) for w in text.replace(',', ' , ').split()]


def _embedder(cfg, W):
    from t2v_b200.clip import FrozenOpenCLIPEmbedder
    e = FrozenOpenCLIPEmbedder(width=cfg.width, heads=cfg.heads, layers=cfg.layers, vocab=cfg.vocab, tokenizer=FakeTokenizer())
    sd = e.model.state_dict()
    sd.update(W)
    e.model.load_state_dict(sd)
    e.model.half().cuda()
    return e


@pytest.mark.gpu
@pytest.mark.parametrize('cfg', [CO.ClipConfig(width=128, heads=2, layers=4, layers_run=3, context=77, vocab=300), CO.ClipConfig()],
                         ids=['narrow', 'ViT-H-14'])
def test_text_tower_vs_oracle(cfg):
    W = UO.make_weights(CO.clip_param_specs(cfg), seed=4)
    Wh = {k: v.half().float() for k, v in W.items()}
    e = _embedder(cfg, W)
    tok = torch.randint(0, cfg.vocab, (2, cfg.context), generator=torch.Generator().manual_seed(2))
    out = e.encode_with_transformer(tok).float().cpu()
    ref = CO.clip_text_forward(Wh, cfg, tok)
    rms = ((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    mx = ((out - ref).abs().max() / ref.abs().max()).item()
    print(f'[parity] clip tower {cfg.width}x{cfg.layers_run}: rel-rms {rms:.3e} max {mx:.3e}')
    assert rms < 3e-3 and mx < 1e-2, (rms, mx)
    assert torch.equal(e.encode_with_transformer(tok).float().cpu(), out)                 # graph replay, deterministic


@pytest.mark.gpu
def test_prompt_chunks_weights_and_padding():
    cfg = CO.ClipConfig(width=128, heads=2, layers=4, layers_run=3, context=77, vocab=300)
    W = UO.make_weights(CO.clip_param_specs(cfg), seed=4)
    Wh = {k: v.half().float() for k, v in W.items()}
    e = _embedder(cfg, W)
    short = 'a cat riding a bike, cinematic'
    long = ' '.join(f'word{i}' for i in range(100))                      # 100 tokens -> two 75-token chunks
    z = e.encode([short, long])
    assert z.shape == (2, 154, 128)
    chunks, count = e.tokenize_line(long)
    assert len(chunks) == 2 and count == 100 and all(len(c.tokens) == 77 for c in chunks)
    assert chunks[0].tokens[0] == 298 and chunks[0].tokens[-1] == 299 and chunks[1].tokens[26] == 299
    # first chunk of the short prompt: tokens after the first <end_of_text> are replaced by the pad id before the transformer
    c0 = e.tokenize_line(short)[0][0]
    toks = torch.tensor([c0.tokens])
    idx = c0.tokens.index(299)
    toks[0, idx + 1:] = 0
    ref = CO.process_tokens(CO.clip_text_forward(Wh, cfg, torch.cat([toks, torch.tensor([chunks[0].tokens])])),
                            [c0.multipliers, chunks[0].multipliers])
    got = z[:, :77].float().cpu()
    assert ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item() < 3e-3
    # emphasis multipliers rescale tokens and restore the mean (clip_hardcode.py:416-420)
    w = [[1.0] * 77, [1.0] * 3 + [1.3] * 10 + [1.0] * 64]
    zw = e.process_tokens([c0.tokens, chunks[0].tokens], w).float().cpu()
    refw = CO.process_tokens(CO.clip_text_forward(Wh, cfg, torch.cat([toks, torch.tensor([chunks[0].tokens])])), w)
    assert ((zw - refw).pow(2).mean().sqrt() / refw.pow(2).mean().sqrt()).item() < 3e-3


def test_prompt_chunking_host_logic_on_cpu(monkeypatch):
    """Pure host logic of FrozenOpenCLIPEmbedder (clip_hardcode.py:146-260, :395-420): 75-token chunks framed by start / end ids,
    BREAK starts a new chunk, padding after the first end token, emphasis multipliers with mean restoration -- checked without a
    GPU by standing the CPU oracle in for the transformer (only here, as the checker)."""
    from t2v_b200.clip import FrozenOpenCLIPEmbedder
    cfg = CO.ClipConfig(width=128, heads=2, layers=4, layers_run=3, context=77, vocab=300)
    with torch.device('meta'):
        e = FrozenOpenCLIPEmbedder(width=cfg.width, heads=cfg.heads, layers=cfg.layers, vocab=cfg.vocab, tokenizer=FakeTokenizer())
    W = UO.make_weights(CO.clip_param_specs(cfg), seed=4)
    seen = {}

    def fake_transformer(tokens):
        seen['tokens'] = tokens.clone()
        return CO.clip_text_forward(W, cfg, tokens)
    monkeypatch.setattr(e, 'encode_with_transformer', fake_transformer)
    long = ' '.join(f'word{i}' for i in range(160))                      # 160 tokens -> 75 + 75 + 10
    chunks, count = e.tokenize_line(long)
    assert count == 160 and len(chunks) == 3 and all(len(c.tokens) == 77 and len(c.multipliers) == 77 for c in chunks)
    assert all(c.tokens[0] == e.id_start and c.tokens[-1] == e.id_end for c in chunks)
    assert chunks[2].tokens[11:] == [e.id_end] * 66 and e.get_target_prompt_token_count(count) == 225
    assert e.get_target_prompt_token_count(0) == 75 and len(e.empty_chunk().tokens) == 77
    monkeypatch.setattr(e, '_parse', lambda line: [['a b', 1.0], ['BREAK', -1], ['c', 1.4]])
    chunks, count = e.tokenize_line('ignored')
    assert len(chunks) == 2 and count == 3 and chunks[1].multipliers[1] == 1.4 and chunks[1].multipliers[2] == 1.0
    monkeypatch.undo()
    monkeypatch.setattr(e, 'encode_with_transformer', fake_transformer)
    z = e.encode(['a cat', long])                                         # batch of prompts with 1 and 3 chunks
    assert z.shape == (2, 3 * 77, 128)
    assert seen['tokens'].shape == (2, 77) and int(seen['tokens'][0, 2:].abs().sum()) == 0      # third chunk of 'a cat': empty chunk, padded
    c0 = e.tokenize_line('a cat')[0][0]
    toks = torch.tensor([c0.tokens])
    toks[0, c0.tokens.index(e.id_end) + 1:] = e.id_pad
    w = [[1.0] * 2 + [1.5] * 3 + [1.0] * 72]
    got = e.process_tokens([c0.tokens], w)
    assert torch.equal(seen['tokens'], toks)
    ref = CO.process_tokens(CO.clip_text_forward(W, cfg, toks), w)
    assert torch.allclose(got, ref, rtol=0, atol=1e-6)

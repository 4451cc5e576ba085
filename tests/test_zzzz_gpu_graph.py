"""DDP_B200_GRAPH=1 (opt-in latency mode: the launch sequence of ddp_sample captured into a CUDA graph and replayed)
must give the very bits of the ordinary launches AND must really replay a graph: the replay / capture counters of the
C ABI (ddp_graph_replays, ddp_graph_captures, ddp_graph_last_fallback) are asserted call by call, so a silent fall-back
to ordinary launches fails the test."""
import pytest
import torch

from oracle import ddp_oracle as O

pytestmark = [pytest.mark.gpu]


def _engine(cfg, W, mode):
    from ddp_b200 import DecodeEngine
    eng = DecodeEngine(task=cfg.task, num_classes=cfg.num_classes, timesteps=cfg.timesteps, bit_scale=cfg.bit_scale,
                       accumulation=cfg.accumulation, min_depth=cfg.min_depth, max_depth=cfg.max_depth, gemm_mode=mode)
    eng.load_state_dict(W)
    return eng


@pytest.mark.parametrize("task,mode", [("seg", "tc_3xf16"), ("seg", "fp32"), ("depth", "tc_3xf16")])
def test_graph_replay_equals_ordinary_launches(monkeypatch, task, mode):
    cfg = O.OracleConfig(task=task, num_classes=19, timesteps=3, randsteps=2, bit_scale=0.01 if task == "seg" else 0.1)
    W = O.make_weights(cfg, seed=21)
    x1, n1 = O.make_inputs(cfg, B=2, h=12, w=20, seed=22)
    x2, n2 = O.make_inputs(cfg, B=2, h=12, w=20, seed=23)
    monkeypatch.delenv("DDP_B200_GRAPH", raising=False)
    plain = _engine(cfg, W, mode)
    want1, want2 = plain.sample(x1.cuda(), n1.cuda()), plain.sample(x2.cuda(), n2.cuda())
    monkeypatch.setenv("DDP_B200_GRAPH", "1")
    eng = _engine(cfg, W, mode)
    x, n = x1.cuda(), n1.cuda()
    out = torch.empty_like(want1)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()               # the legacy default stream cannot be captured (graph mode bypasses itself there)
    with torch.cuda.stream(side):
        for i in range(4):                   # 1st call: ordinary launches, 2nd: capture + launch, 3rd / 4th: replay
            out.zero_()
            eng.sample(x, n, out=out)
            side.synchronize()
            assert torch.equal(out, want1)
            assert eng.graph_replays == (0, 1, 2, 3)[i], eng.graph_last_fallback
            assert eng.graph_captures == (0, 1, 1, 1)[i], eng.graph_last_fallback
            assert (eng.graph_last_fallback == "") == (i > 0)
        assert eng.last_launch_count == plain.last_launch_count
        x.copy_(x2.cuda()); n.copy_(n2.cuda())   # new data through the same buffers: the replay reads them
        eng.sample(x, n, out=out)
        side.synchronize()
        assert torch.equal(out, want2)
        assert eng.graph_replays == 4 and eng.graph_captures == 1
        other = torch.empty_like(out)        # a different output buffer: ordinary launches again, then a new graph
        for i in range(3):
            eng.sample(x, n, out=other)
            side.synchronize()
            assert torch.equal(other, want2)
            assert eng.graph_replays == 4 + i and eng.graph_captures == 1 + (i > 0), eng.graph_last_fallback
        eng.plan(1, 2, 12, 20)               # a new plan invalidates the graph
        got = eng.sample(x[:1].contiguous(), n[:1].contiguous())
        side.synchronize()
    assert torch.equal(got, plain.sample(x[:1].contiguous(), n[:1].contiguous()))
    # on the default stream graph mode steps aside and the ordinary launches run
    before = eng.graph_replays
    assert torch.equal(eng.sample(x[:1].contiguous(), n[:1].contiguous()), got)
    assert eng.graph_replays == before and "default stream" in eng.graph_last_fallback

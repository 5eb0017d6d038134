"""Host logic of the reference's progress loop (src/renderer.rs:205-251) in the Python mirror, with a stub in place
of the device context: stop conditions (max sampling, time limit with the x1.1 prediction) and interval images."""
import numpy as np
import pytest


class StubCtx:
    def __init__(self):
        self.resolved = []

    def resolve(self, sampling, out=None):
        self.resolved.append(sampling)
        if out is not None:
            out[:] = sampling % 256
        return out


class Clock:
    def __init__(self, t=1000.0):
        self.t = t

    def __call__(self):
        return self.t


@pytest.fixture
def clock(monkeypatch, hr):
    c = Clock()
    monkeypatch.setattr(hr.time, "time", c)
    return c


def test_stops_at_max_sampling(hr, clock):
    r = hr.PathTracingRenderer(3, time_limit_sec=1e9, report_interval_sec=1e9)
    ctx, img = StubCtx(), np.zeros((2, 2, 3), np.uint8)
    clock.t += 1.0
    assert r.report_progress(ctx, 1, img) is False
    clock.t += 1.0
    assert r.report_progress(ctx, 2, img) is False
    assert ctx.resolved == []                       # no interval image yet, nothing resolved
    clock.t += 1.0
    assert r.report_progress(ctx, 3, img) is True   # sampling >= max_sampling: final image, stop
    assert ctx.resolved == [3] and int(img[0, 0, 0]) == 3


def test_time_limit_uses_the_last_pass_times_1_1(hr, clock):
    r = hr.PathTracingRenderer(1000, time_limit_sec=10.0, report_interval_sec=1e9)
    ctx, img = StubCtx(), np.zeros((1, 1, 3), np.uint8)
    clock.t += 4.0                                   # used 4.0, last pass 4.0: 4 + 4.4 = 8.4 <= 10 -> go on
    assert r.report_progress(ctx, 1, img) is False
    clock.t += 2.0                                   # used 6.0, last pass 2.0: 6 + 2.2 = 8.2 <= 10 -> go on
    assert r.report_progress(ctx, 2, img) is False
    clock.t += 2.5                                   # used 8.5, last pass 2.5: 8.5 + 2.75 = 11.25 > 10 -> stop now
    assert r.report_progress(ctx, 3, img) is True
    assert ctx.resolved == [3]


def test_interval_images_are_numbered(hr, clock):
    saved = []
    r = hr.PathTracingRenderer(1000, time_limit_sec=1e9, report_interval_sec=15.0, save_progress=lambda name, img: saved.append((name, int(img[0, 0, 0]))))
    ctx, img = StubCtx(), np.zeros((1, 1, 3), np.uint8)
    for sampling, dt in ((1, 10.0), (2, 10.0), (3, 10.0), (4, 10.0)):
        clock.t += dt
        assert r.report_progress(ctx, sampling, img) is False
    # 15 s elapsed after pass 2 (t = 20) -> 000.png, then again after pass 4 (t = 40) -> 001.png
    assert saved == [("000.png", 2), ("001.png", 4)]
    assert r.report_image_counter == 2


def test_debug_renderer_is_one_pass(hr):
    d = hr.DebugRenderer(hr.MODE_DEBUG_NORMAL)
    ctx, img = StubCtx(), np.zeros((1, 1, 3), np.uint8)
    assert d.max_sampling() == 1 and d.report_progress(ctx, 1, img) is True and ctx.resolved == [1]

"""Per-op CUDA-event timing for bench.py's roofline bookkeeping (off by default; zero cost when disabled).

Every C-ABI call of the ops / spconv layers is bracketed by `with profiler.record(name, algorithmic_bytes)`.
`algorithmic_bytes` follows SURVEY.md 8(d) / DESIGN.md "Algorithmic bytes"."""
import torch

_enabled = False
_records = []
_steps = 0


def enable():
    global _enabled, _steps
    _enabled = True
    _steps += 1


def disable():
    global _enabled
    _enabled = False


def reset():
    global _records, _steps
    _records = []
    _steps = 0


class record(object):
    __slots__ = ('name', 'nbytes', 'e0', 'e1')

    def __init__(self, name, nbytes=0):
        self.name = name
        self.nbytes = nbytes

    def __enter__(self):
        if _enabled:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if _enabled:
            self.e1.record()
            _records.append((self.name, self.nbytes, self.e0, self.e1))


def summary():
    torch.cuda.synchronize()
    steps = max(_steps, 1)
    agg = {}
    for name, nbytes, e0, e1 in _records:
        a = agg.setdefault(name, [0.0, 0.0, 0])
        a[0] += e0.elapsed_time(e1)
        a[1] += float(nbytes() if callable(nbytes) else nbytes)
        a[2] += 1
    total_ms = sum(a[0] for a in agg.values()) or 1e-9
    by = {}
    for name, (ms, nb, cnt) in agg.items():
        by[name] = dict(ms_per_step=ms / steps, launches_per_step=cnt / steps, bytes_per_step=nb / steps,
                        gbs=(nb / 1e9) / (ms / 1e3) if ms > 0 else 0.0, avg_us=ms / cnt * 1e3,
                        share=ms / total_ms)
    dom_name = max(by, key=lambda k: by[k]['ms_per_step']) if by else None
    dom = dict(name=dom_name, **by[dom_name]) if dom_name else dict(name=None, gbs=0.0, launches_per_step=0,
                                                                    avg_us=0.0, share=0.0, bytes_per_step=0.0)
    return dict(dominant=dom, by_kernel={k: {kk: round(vv, 4) for kk, vv in v.items()} for k, v in by.items()},
                stage_ms={k: round(v['ms_per_step'], 4) for k, v in by.items()})

"""Build the UNMODIFIED reference ops extension into oracle/_ref/ (test infrastructure only).

Compiles /root/reference/softgroup/ops/src/{softgroup_api.cpp,softgroup_ops.cpp,cuda.cu}
where they lie (no sources are copied) with an include-path shim for the missing
sparsehash header. Output: oracle/_ref/sg_ref_ops*.so (git-ignored, travels to the GPU box).
The CPU entry points (voxelize_idx, bfs_cluster, build_and_export_octree) run anywhere;
the CUDA entry points (ballquery_batch_p, voxelize_fp, sec_*, ...) run on the GPU box and
serve as a second oracle there.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import the result.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/softgroup/ops/src'
OUT = os.path.join(HERE, '_ref')
NAME = 'sg_ref_ops'


def built_path():
    if not os.path.isdir(OUT):
        return None
    for f in os.listdir(OUT):
        if f.startswith(NAME) and f.endswith('.so'):
            return os.path.join(OUT, f)
    return None


def build(verbose=False):
    """Returns path of the built .so, or None when the reference tree is absent."""
    p = built_path()
    if p is not None:
        return p
    if not os.path.isdir(REF_SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    os.environ.setdefault('MAX_JOBS', '8')
    from torch.utils.cpp_extension import load
    load(
        name=NAME,
        sources=[os.path.join(REF_SRC, s) for s in ('softgroup_api.cpp', 'softgroup_ops.cpp', 'cuda.cu')],
        extra_include_paths=[os.path.join(HERE, 'shim')],
        extra_cflags=['-O2', '-w'],
        extra_cuda_cflags=['-O2', '-w'],
        build_directory=OUT,
        verbose=verbose,
        is_python_module=True)
    return built_path()


def load_ref():
    """Import the built reference extension (None if unavailable)."""
    p = built_path()
    if p is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv))

"""CPU oracle for the resselt hot path (forward of the in-scope SR architectures).

TEST INFRASTRUCTURE ONLY.  Nothing in ``resselt_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may.

What it is: a functional restatement (``torch.nn.functional`` on CPU tensors, fp32 or fp64) of
the reference's ``nn.Module.forward`` for each architecture, taking a plain state dict.  The
reference's arithmetic itself lives in a third-party dependency, PyTorch ATen (torch>=2.6,
/root/reference/pyproject.toml:11); this oracle calls the same ATen CPU ops, so it differs from
the reference only by the order of a few fp32 reductions (Conv3XC merge in closed form).

Parity pin: the reference has no tests/golden vectors of its own, so the oracle is
pinned against outputs of the reference itself, generated in the build container by
``oracle/make_golden.py`` (which imports /root/reference) and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` replays them without the reference.
"""
from .sr_forward import (  # noqa: F401
    compact_forward,
    dat_forward,
    esrgan_forward,
    forward_by_name,
    plksr_forward,
    realplksr_forward,
    rtmosr_forward,
    gaterv3_forward,
    span_forward,
    spanplus_forward,
    spanpp_forward,
    swinir_forward,
)

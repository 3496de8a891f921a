"""TEST INFRASTRUCTURE -- seeded input builders shared by `make_golden.py` and the tests, so the
fixtures under tests/golden/ only need to hold the reference's OUTPUTS (torch's CPU generator is
deterministic for a given torch build; the GPU box runs the same image)."""
import torch


def crandn(seed, *shape):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) + 1j * torch.randn(*shape, generator=g)).to(torch.complex64)


def rrand(seed, *shape):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed))


def rrandn(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def phantom(seed, *shape):
    """magnitude U[0,1) with a random phase -- stands in for `add_phase` (helpers/load_data.py:372-387)."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * torch.exp(1j * torch.randn(*shape, generator=g))).to(torch.complex64)

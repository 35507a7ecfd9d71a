import torch


def gen_f32(n_elems, device="cuda", f64=False):
    """BASELINE C2/C3/C4 field generated on the device (float32, or float64 viewed as bytes)."""
    n = n_elems // 2 if f64 else n_elems
    i = torch.arange(n, device=device, dtype=torch.float64)
    g = torch.Generator(device=device); g.manual_seed(0xB200)
    u = torch.rand(n, device=device, generator=g, dtype=torch.float64) * 2 - 1
    x = torch.sin(2 * torch.pi * i / 4096) + 0.25 * torch.sin(2 * torch.pi * i / 333.3) + 1e-3 * u
    return (x if f64 else x.to(torch.float32)).view(torch.uint8)

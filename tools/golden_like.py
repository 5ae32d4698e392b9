def rel(y, ref):
    y, ref = y.double(), ref.double()
    return ((y - ref).abs().max() / ref.abs().max().clamp(min=1e-30)).item()

"""Import shim: AutoencoderKL is outside the north-star path; the oracle never builds it."""


class AutoencoderKL:
    @classmethod
    def from_pretrained(cls, *a, **kw):
        raise RuntimeError("diffusers is not available; the VAE is out of the measured path")

"""Entropy models of the HESIC path on the hesic_b200 kernels.

Mirrors compressai/entropy_models/entropy_models.py: ``EntropyModel`` (:56-239),
``EntropyBottleneck`` (:242-430), ``GaussianConditional`` (:433-562) and the authors'
``GaussianMixtureConditional`` (:566-710) -- same constructor arguments, buffers, ``state_dict``
keys, return values and ValueErrors.  ``forward`` (eval mode: quantise + likelihood + clamp) is one
fused sm_100a kernel per call; ``compress`` prepares int32 symbols / CDF indexes on the device
(bit-exact with the reference) and feeds the host rANS coder with flat arrays; ``update`` builds the
CDF tables once per model on the host, as the reference does.  Training-mode forward (additive
uniform noise) is outside the inference hot path and raises NotImplementedError.
"""
import numpy as np
import scipy.stats
import torch
import torch.nn as nn
import torch.nn.functional as TF

from compressai.ops import LowerBound
from hesic_b200 import _capi as _C
from hesic_b200 import functional as _F


class _EntropyCoder:
    """Proxy to the actual coder (entropy_models.py:13-42)."""

    def __init__(self, method):
        if not isinstance(method, str):
            raise ValueError(f'Invalid method type "{type(method)}"')
        from compressai import available_entropy_coders
        if method not in available_entropy_coders():
            methods = ", ".join(available_entropy_coders())
            raise ValueError(f'Unknown entropy coder "{method}" (available: {methods})')
        if method == "ans":
            from compressai import ans
            self._encoder, self._decoder = ans.RansEncoder(), ans.RansDecoder()
        elif method == "rangecoder":
            import range_coder
            self._encoder, self._decoder = range_coder.RangeEncoder(), range_coder.RangeDecoder()

    def encode_with_indexes(self, *args, **kwargs):
        return self._encoder.encode_with_indexes(*args, **kwargs)

    def decode_with_indexes(self, *args, **kwargs):
        return self._decoder.decode_with_indexes(*args, **kwargs)


def default_entropy_coder():
    from compressai import get_entropy_coder
    return get_entropy_coder()


def pmf_to_quantized_cdf(pmf, precision=16):
    cdf = _F.pmf_to_quantized_cdf(pmf.detach().cpu().numpy(), precision)
    return torch.from_numpy(cdf.astype(np.int32))


class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound=1e-9, entropy_coder=None, entropy_coder_precision=16):
        super().__init__()
        if entropy_coder is None:
            entropy_coder = default_entropy_coder()
        self.entropy_coder = _EntropyCoder(entropy_coder)
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.likelihood_bound = float(likelihood_bound)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    def forward(self, *args):
        raise NotImplementedError()

    def _lik_bound(self):
        return self.likelihood_bound if self.use_likelihood_bound else 0.0

    def _quantize(self, inputs, mode, means=None):
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
        _C.require_cuda(inputs)
        if mode == "symbols":
            return _F.prepare_symbols(inputs, self._expand_means(inputs, means)).reshape(inputs.shape)
        if means is None:
            return _F.round_half_even(inputs)
        m = means.expand_as(inputs).contiguous()
        return _F.round_half_even(inputs - m) + m

    @staticmethod
    def _expand_means(inputs, means):
        if means is None:
            return None
        if means.numel() == inputs.size(1) or means.shape == inputs.shape:
            return means
        return means.expand_as(inputs).contiguous()

    @staticmethod
    def _dequantize(inputs, means=None):
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.float()
        return outputs

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        pmf, tail_mass = pmf.detach().cpu(), tail_mass.detach().cpu()
        for i, p in enumerate(pmf):
            prob = torch.cat((p[:pmf_length[i]], tail_mass[i]), dim=0)
            _cdf = pmf_to_quantized_cdf(prob, self.entropy_coder_precision)
            cdf[i, :_cdf.size(0)] = _cdf
        return cdf

    def _check_cdf_size(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if len(self._quantized_cdf.size()) != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def _check_offsets_size(self):
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if len(self._offset.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")

    def _check_cdf_length(self):
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if len(self._cdf_length.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    def _host_tables(self):
        return (self._quantized_cdf.detach().cpu().numpy().astype(np.int32),
                self._cdf_length.detach().cpu().reshape(-1).numpy().astype(np.int32),
                self._offset.detach().cpu().reshape(-1).numpy().astype(np.int32))

    def compress(self, inputs, indexes, means=None):
        """Symbols and indexes are prepared on the device; one D2H copy each, then the host coder."""
        if len(inputs.size()) != 4:
            raise ValueError("Invalid `inputs` size. Expected a 4-D tensor.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        symbols = self._quantize(inputs, "symbols", means)
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        sym = symbols.reshape(symbols.size(0), -1).cpu().numpy()
        idx = indexes.reshape(indexes.size(0), -1).int().cpu().numpy()
        cdf, length, offset = self._host_tables()
        return [self.entropy_coder.encode_with_indexes(sym[i], idx[i], cdf, length, offset) for i in range(sym.shape[0])]

    def decompress(self, strings, indexes, means=None):
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if len(indexes.size()) != 4:
            raise ValueError("Invalid `indexes` size. Expected a 4-D tensor.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        if means is not None:
            if means.size()[:-2] != indexes.size()[:-2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size() and (means.size(2) != 1 or means.size(3) != 1):
                raise ValueError("Invalid means parameters")
        cdf, length, offset = self._host_tables()
        idx = indexes.reshape(indexes.size(0), -1).int().cpu().numpy()
        vals = np.stack([np.asarray(self.entropy_coder.decode_with_indexes(s, idx[i], cdf, length, offset), dtype=np.int32)
                         for i, s in enumerate(strings)])
        outputs = torch.from_numpy(vals).reshape(indexes.size()).to(self._quantized_cdf.device)
        return self._dequantize(outputs, means)


class EntropyBottleneck(EntropyModel):
    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        if self.filters != (3, 3, 3, 3):
            # the only configuration on the path (newnet1.py:42-45 builds EntropyBottleneck(channels) with the default);
            # refuse at construction rather than at the first forward
            raise NotImplementedError(f"hesic_b200 EntropyBottleneck supports filters=(3, 3, 3, 3) only (the reference's "
                                      f"default and the only value HSIC / DSIC use), got {self.filters}")
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)

        self._biases = nn.ParameterList()
        self._factors = nn.ParameterList()
        self._matrices = nn.ParameterList()
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            self._matrices.append(nn.Parameter(torch.full((self.channels, filters[i + 1], filters[i]), float(init))))
            self._biases.append(nn.Parameter(torch.empty(self.channels, filters[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self._factors.append(nn.Parameter(torch.zeros(self.channels, filters[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.Tensor([-self.init_scale, 0, self.init_scale]).repeat(self.channels, 1, 1))
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))
        self._packed = None
        self._packed_key = None

    def _medians(self):
        return self.quantiles[:, :, 1:2]

    def hesic_params(self):
        """Device table of pre-activated parameters (softplus(M), tanh(f), median), rebuilt on change."""
        if self.filters != (3, 3, 3, 3):
            raise NotImplementedError("hesic_b200 EntropyBottleneck kernel is specialised for filters=(3,3,3,3)")
        ts = list(self._matrices) + list(self._biases) + list(self._factors) + [self.quantiles]
        key = tuple((t.data_ptr(), t._version, t.device.index) for t in ts)
        if key != self._packed_key:
            self._packed = _F.eb_pack(list(self._matrices), list(self._biases), list(self._factors), self.quantiles)
            self._packed_key = key
        return self._packed

    def update(self, force=False):
        if self._offset.numel() > 0 and not force:
            return
        with torch.no_grad():
            medians = self.quantiles[:, 0, 1]
            minima = torch.clamp(torch.ceil(medians - self.quantiles[:, 0, 0]).int(), min=0)
            maxima = torch.clamp(torch.ceil(self.quantiles[:, 0, 2] - medians).int(), min=0)
            self._offset = -minima
            pmf_start = medians - minima
            pmf_length = maxima + minima + 1
            max_length = pmf_length.max()
            samples = torch.arange(max_length, device=pmf_start.device)
            samples = samples[None, :] + pmf_start[:, None, None]
            lower = self._logits_cumulative(samples - 0.5, stop_gradient=True)
            upper = self._logits_cumulative(samples + 0.5, stop_gradient=True)
            sign = -torch.sign(lower + upper)
            pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
            tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
            quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
            self._quantized_cdf = quantized_cdf.to(self.quantiles.device)
            self._cdf_length = pmf_length + 2

    def loss(self):
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def _logits_cumulative(self, inputs, stop_gradient):
        """Table-building / aux-loss helper (runs once per model, plain torch, any device)."""
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = self._matrices[i].detach() if stop_gradient else self._matrices[i]
            logits = torch.matmul(TF.softplus(matrix), logits)
            logits = logits + (self._biases[i].detach() if stop_gradient else self._biases[i])
            if i < len(self._factors):
                factor = self._factors[i].detach() if stop_gradient else self._factors[i]
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def forward(self, x):
        if self.training:
            raise NotImplementedError("hesic_b200: training-mode (noise) forward is outside the inference hot path")
        _C.require_cuda(x)
        return _F.entropy_bottleneck(x, self.hesic_params(), self._lik_bound())

    @staticmethod
    def _build_indexes(size, device=None):
        N, Cn, H, W = size
        if device is not None and torch.device(device).type == "cuda":
            return _F.build_indexes_channel((N, Cn, H, W), device)
        return torch.arange(Cn).view(1, -1, 1, 1).int().repeat(N, 1, H, W)

    def compress(self, x):
        indexes = self._build_indexes(x.size(), x.device)
        medians = self._medians().detach().view(1, -1, 1, 1)
        return super().compress(x, indexes, medians)

    def decompress(self, strings, size):
        output_size = (len(strings), self._quantized_cdf.size(0), size[0], size[1])
        indexes = self._build_indexes(output_size)
        medians = self._medians().detach().view(1, -1, 1, 1)
        return super().decompress(strings, indexes, medians)


class _GaussianBase(EntropyModel):
    """Shared scale-table machinery of GaussianConditional / GaussianMixtureConditional."""

    def _init_scales(self, scale_table, scale_bound, tail_mass):
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            self.lower_bound_scale = LowerBound(self.scale_table[0])
        elif scale_bound > 0:
            self.lower_bound_scale = LowerBound(scale_bound)
        else:
            raise ValueError("Invalid parameters")

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    def _standardized_cumulative(self, inputs):
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        return scipy.stats.norm.ppf(quantile)

    def _scale_bound_value(self):
        # cached on the host: reading the buffer every forward would be a device->host sync
        key = (self.lower_bound_scale.bound.data_ptr(), self.lower_bound_scale.bound._version)
        if getattr(self, "_bound_cache", (None, None))[0] != key:
            self._bound_cache = (key, float(self.lower_bound_scale.bound.item()))
        return self._bound_cache[1]

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return
        self.scale_table = self._prepare_scale_table(scale_table).to(self.scale_bound.device
                                                                     if self.scale_bound is not None else "cpu")
        self.update()

    def update(self):
        table = self.scale_table.detach().cpu()
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = torch.max(pmf_length).item()
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        samples_scale = table.unsqueeze(1).float()
        upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
        lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        dev = self.scale_table.device
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(dev)
        self._offset = (-pmf_center).to(dev)
        self._cdf_length = (pmf_length + 2).to(dev)

    def build_indexes(self, scales):
        """63 - #{s in table[:-1] : max(scale, bound) <= s}, one kernel (entropy_models.py:556-562)."""
        _C.require_cuda(scales)
        return _F.build_indexes_scale(scales, self.scale_table.to(scales.device), self._scale_bound_value())


class GaussianConditional(_GaussianBase):
    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        self._init_scales(scale_table, scale_bound, tail_mass)

    def _likelihood(self, inputs, scales, means=None):
        """Likelihood of already-quantised inputs (entropy_models.py:528-544)."""
        _C.require_cuda(inputs, scales)
        # the fused kernel quantises first; for already-quantised inputs that is the identity
        return _F.gaussian_conditional(inputs, scales, means, self._scale_bound_value(), 0.0)[1]

    def forward(self, inputs, scales, means=None):
        if self.training:
            raise NotImplementedError("hesic_b200: training-mode (noise) forward is outside the inference hot path")
        _C.require_cuda(inputs, scales)
        return _F.gaussian_conditional(inputs, scales, means, self._scale_bound_value(), self._lik_bound())


class GaussianMixtureConditional(_GaussianBase):
    """The authors' K-component mixture (entropy_models.py:566-710).  Channel k*M+m of
    scales/means/weights is component k of latent channel m; quantisation ignores the means (:697)."""

    def __init__(self, K, scale_table=None, mean_table=None, weight_table=None, *args, scale_bound=0.11,
                 tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        self.K = K
        self._init_scales(scale_table, scale_bound, tail_mass)

    def _likelihood(self, inputs, scales, means=None, weights=None):
        _C.require_cuda(inputs, scales, means, weights)
        return _F.gaussian_mixture_conditional(inputs, scales, means, weights, self.K, self._scale_bound_value(), 0.0)[1]

    def forward(self, inputs, scales, means=None, weights=None):
        if self.training:
            raise NotImplementedError("hesic_b200: training-mode (noise) forward is outside the inference hot path")
        _C.require_cuda(inputs, scales, means, weights)
        return _F.gaussian_mixture_conditional(inputs, scales, means, weights, self.K, self._scale_bound_value(),
                                               self._lik_bound())

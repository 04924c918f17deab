"""Attention-module integration hooks.

The three attention modules the reference patches all run the same sequence between their qkv
projection and their output projection:
    workloads/deit/scripts/main.py:85-157                       QuantizedAttention.forward
    workloads/DiT/models.py:154-230                             Attention.forward
    workloads/PixArt/models/MX_transformer_block.py:624-717     MXSelfAttention.forward
``PrunedAttentionCore`` is that sequence (lines 101-152 of the DeiT file) as one call into
libmxprune; the three shims keep the reference constructors' / ``set_config`` argument names so a
maintainer can swap the body of ``forward`` (INTEGRATION.md shows the patch).  The qkv / output
projections are ``MxLinear`` (the reference's ``mx.Linear`` forward on the tcgen05 GEMM) when the shim is
built with ``mx_quant=True``, else whatever the host model uses.

Implemented: mx_quant && top_k && approx/ex_pred with pred_mode "ex_pred" (the pruned hot path),
"partial_Q", "partial_K", "MXINT4", "two_step_leading_ones", "true_ex" or "ELSA" (with the caller's orthogonal matrix);
mx_quant && top_k && !approx (top-k of the true scores); and mx_quant && !top_k (dense MXINT8 attention, what the
reference runs in the last block of each model - same kernels with every key kept).  mx_quant=False (the reference's
unquantized torch path) raises - no silent fallback.
"""
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .specs import resolve_linear_specs, resolve_specs


class PrunedAttentionCore(nn.Module):
    """q, k, v (B,H,N,hd) fp32 views -> x (B,N,H*hd), ready for the output projection."""

    def __init__(self, mx_specs, k: int, scale: Optional[float] = None, pred_mode: str = "ex_pred",
                 orthogonal_matrix: Optional[torch.Tensor] = None):
        super().__init__()
        resolve_specs(mx_specs)
        self.mx_specs = mx_specs
        self.k = int(k)
        self.scale = scale
        self.pred_mode = pred_mode
        if pred_mode == "ELSA" and orthogonal_matrix is None:
            raise ValueError("pred_mode='ELSA' needs orthogonal_matrix (workloads/deit/scripts/main.py:119-121)")
        self.orthogonal_matrix = orthogonal_matrix
        # --anal (main.py:134-136): "Average chosen k" = funcs/analysis.py total_chosen_k, from the kept-key bitmask;
        # the per-timestep diff_idx text dumps (main.py:137-143) are not produced
        self.anal = False
        self.avg_chosen_k = None

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                key_bias: Optional[torch.Tensor] = None, dense: bool = False,
                pred_mode: Optional[str] = None) -> torch.Tensor:
        """dense / pred_mode: per-call overrides for the reference's exclude_timesteps steps."""
        B, H, N, hd = q.shape
        buf = torch.empty((B, N, H, hd), dtype=torch.float32, device=q.device)
        # write straight into (B,N,H,hd): the reference's x.transpose(1,2).reshape(B,N,C) is free
        # k <= 0: dense MXINT8 attention (the reference's top_k=False blocks) = every key kept
        top_k = self.k if (self.k > 0 and not dense) else k.shape[2]
        mode = pred_mode or self.pred_mode
        res = ops.pruned_attention(q, k, v, self.mx_specs, top_k, scale=self.scale, out=buf.permute(0, 2, 1, 3),
                                   key_bias=key_bias, pred_mode=mode, return_mask=self.anal,
                                   orthogonal_matrix=self.orthogonal_matrix if mode == "ELSA" else None)
        if self.anal:
            from .analysis import coverage_rate
            self.avg_chosen_k = coverage_rate(res[1])
            print(f"Average chosen k: {self.avg_chosen_k:.3f}")         # as main.py:136
        return buf.reshape(B, N, H * hd)


def _require_hot_path(mx_quant, top_k, approx, pred_mode, where) -> str:
    """Validate the reference's flag combination; returns the ops.PRED_MODES key that ranks the keys
    (workloads/deit/scripts/main.py:104-131: approx_flag picks the predictor, else top-k of the true scores)."""
    if mx_quant and not top_k:
        return "ex_pred"    # dense MXINT8 attention (deit main.py:282-296: the last block runs top_k=False)
    if not mx_quant:
        raise NotImplementedError(f"{where}: mx_quant=False (the fp32 attention of the host model) is not on the "
                                  "B200 path and there is no fallback")
    if not approx:
        return "exact"
    if pred_mode not in ("ex_pred", "partial_Q", "partial_K", "MXINT4", "two_step_leading_ones", "true_ex", "ELSA"):
        raise NotImplementedError(
            f"{where}: pred_mode={pred_mode!r} is not built (built: 'ex_pred', 'partial_Q', 'partial_K', 'MXINT4', "
            "'two_step_leading_ones', 'true_ex', 'ELSA' and approx=False) "
            "(SURVEY.md 8f3) and there is no fallback")
    return pred_mode


def to_mx_linear(lin: nn.Linear, mx_specs) -> "MxLinear":
    """nn.Linear -> MxLinear sharing the parameters (what apply_quantization_to_deit /
    MXBasicTransformerBlock.set_config do with mx.Linear: deit main.py:231-318,
    MX_transformer_block.py:344-362)."""
    if isinstance(lin, MxLinear):
        return lin
    m = MxLinear(lin.in_features, lin.out_features, bias=lin.bias is not None, mx_specs=mx_specs)
    m.weight = lin.weight
    if lin.bias is not None:
        m.bias = lin.bias
    return m


class QuantizedAttention(nn.Module):
    """DeiT shim - constructor mirrors workloads/deit/scripts/main.py:42."""

    def __init__(self, orig_attn, mx_quant=False, mx_specs=None, top_k=True, k=20, approx_flag=True,
                 pred_mode="ex_pred", anal=False, file_name_dict=None, block_idx=None, orthogonal_matrix=None):
        super().__init__()
        mode = _require_hot_path(mx_quant, top_k, approx_flag, pred_mode, "QuantizedAttention")
        self.num_heads = orig_attn.num_heads
        self.scale = orig_attn.scale
        # the reference swaps the projections for mx.Linear as well (main.py:262-281)
        self.qkv, self.proj = to_mx_linear(orig_attn.qkv, mx_specs), to_mx_linear(orig_attn.proj, mx_specs)
        self.proj_drop = getattr(orig_attn, "proj_drop", nn.Identity())
        self.block_idx = block_idx
        self.current_timestep = 0
        self.core = PrunedAttentionCore(mx_specs, k if top_k else 0, scale=self.scale, pred_mode=mode,
                                        orthogonal_matrix=orthogonal_matrix)
        self.core.anal = bool(anal) and bool(top_k)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        x = self.core(qkv[0], qkv[1], qkv[2])         # strided views of the fused buffer, no copies
        x = self.proj_drop(self.proj(x))
        self.current_timestep += 1
        return x


class QuantizedMlp(nn.Module):
    """DeiT MLP with MX Linear layers - mirrors workloads/deit/scripts/main.py:159-196 (fc1 -> act -> fc2 -> drop;
    the activation stays the host model's own module, as in the reference)."""

    def __init__(self, orig_mlp, mx_specs=None):
        super().__init__()
        self.mx_specs = mx_specs
        self.act = orig_mlp.act
        self.drop = nn.Dropout(getattr(getattr(orig_mlp, "drop", None), "p", 0.0))
        self.fc1, self.fc2 = to_mx_linear(orig_mlp.fc1, mx_specs), to_mx_linear(orig_mlp.fc2, mx_specs)

    def forward(self, x):
        return self.drop(self.fc2(self.act(self.fc1(x))))


def apply_quantization_to_deit(model, config, mx_quant=True, top_k=True, k=20, approx_flag=True, pred_mode="ex_pred",
                               anal=False, file_name_dict=None, exclude_blocks=(), exclude_block_type="ex_pred",
                               orthogonal_matrix=None, dense_blocks=(11,)):
    """Swap the attention / MLP modules of a timm-style DeiT (``model.blocks[i].attn`` / ``.mlp``) for the shims, with
    the reference's block policy (workloads/deit/scripts/main.py:231-318): the blocks in ``dense_blocks`` - index 11, as
    the reference hard-codes it (:266,:281,:296) - run dense MXINT8 attention (top_k=False), blocks in
    ``exclude_blocks`` use ``exclude_block_type`` as their pred_mode, every other listed block uses ``pred_mode``."""
    block_indices = config.get('blocks', [])
    components = config.get('components', ['attn', 'ffn'])
    mx_specs = config.get('mx_specs')
    if mx_specs is None:
        raise ValueError("config['mx_specs'] is required (the reference's default dict lacks keys this path validates)")
    dense_blocks = set(dense_blocks)
    for idx in block_indices:
        if idx >= len(model.blocks):
            continue
        block = model.blocks[idx]
        if 'attn' in components:
            dense, excluded = idx in dense_blocks, idx in exclude_blocks
            block.attn = QuantizedAttention(
                orig_attn=block.attn, mx_quant=mx_quant, mx_specs=mx_specs, top_k=top_k and not dense, k=k,
                approx_flag=approx_flag, pred_mode=exclude_block_type if (dense or excluded) else pred_mode,
                anal=anal, file_name_dict=file_name_dict, block_idx=idx, orthogonal_matrix=orthogonal_matrix)
        if 'ffn' in components:
            block.mlp = QuantizedMlp(orig_mlp=block.mlp, mx_specs=mx_specs)
    return model


class Attention(nn.Module):
    """DiT shim - constructor mirrors workloads/DiT/models.py:105-126."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_norm=False, proj_bias=True, attn_drop=0.,
                 proj_drop=0., norm_layer=nn.LayerNorm, mx_quant=False, mx_specs=None, top_k=False, k=20,
                 ex_pred=False, pred_mode="ex_pred", anal=False, file_name_dict=None, block_idx=None,
                 exclude_timesteps=None, orthogonal_matrix=None):
        super().__init__()
        assert dim % num_heads == 0, 'dim should be divisible by num_heads'
        mode = _require_hot_path(mx_quant, top_k, ex_pred, pred_mode, "Attention")
        # steps listed here run dense attention (models.py:172: `top_k and current_timestep not in exclude_timesteps`)
        self.exclude_timesteps = set(exclude_timesteps or ())
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = MxLinear(dim, dim * 3, bias=qkv_bias, mx_specs=mx_specs)        # models.py:129 (mx Linear)
        self.q_norm = norm_layer(self.head_dim) if qk_norm else nn.Identity()
        self.k_norm = norm_layer(self.head_dim) if qk_norm else nn.Identity()
        self.proj = MxLinear(dim, dim, bias=proj_bias, mx_specs=mx_specs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.block_idx = block_idx
        self.current_timestep = 0
        self.core = PrunedAttentionCore(mx_specs, k if top_k else 0, scale=self.scale, pred_mode=mode,
                                        orthogonal_matrix=orthogonal_matrix)
        self.core.anal = bool(anal) and bool(top_k)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        q, k = self.q_norm(q), self.k_norm(k)
        x = self.core(q, k, v, dense=self.current_timestep in self.exclude_timesteps)
        x = self.proj_drop(self.proj(x))
        self.current_timestep += 1
        return x


class MXSelfAttention(nn.Module):
    """PixArt-alpha shim - mirrors workloads/PixArt/models/MX_transformer_block.py:567-717."""

    def __init__(self, dim, num_heads: int, has_bias: bool = True):
        super().__init__()
        self.dim, self.num_heads, self.head_dim, self.has_bias = dim, num_heads, dim // num_heads, has_bias
        self.to_q = nn.Linear(dim, dim, bias=has_bias)
        self.to_k = nn.Linear(dim, dim, bias=has_bias)
        self.to_v = nn.Linear(dim, dim, bias=has_bias)
        self.to_out = nn.Sequential(nn.Linear(dim, dim, bias=has_bias), nn.Dropout(p=0.0))
        self.core = None
        self.current_timestep = 0

    def set_config(self, mx_quant=False, mx_specs=None, top_k=False, k=20, ex_pred=False, exclude_timesteps=None,
                   pred_mode="ex_pred", block_idx=None, anal=False, file_name_dict=None, orthogonal_matrix=None):
        mode = _require_hot_path(mx_quant, top_k, ex_pred, pred_mode, "MXSelfAttention.set_config")
        # self-attention: listed steps run dense (MX_transformer_block.py:656); cross-attention: listed steps rank
        # on the true scores instead of the predictor's (:806, else-branch :845-848)
        self.exclude_timesteps = set(exclude_timesteps or ())
        self.block_idx = block_idx
        # the block's set_config swaps nn.Linear -> mx.Linear (MX_transformer_block.py:344-362)
        self.to_q, self.to_k, self.to_v = (to_mx_linear(m, mx_specs) for m in (self.to_q, self.to_k, self.to_v))
        self.to_out[0] = to_mx_linear(self.to_out[0], mx_specs)
        # reference: scale_factor = 1 / math.sqrt(q.size(-1)) applied as an fp32 scalar (:647-653)
        self.core = PrunedAttentionCore(mx_specs, k if top_k else 0, scale=1.0 / (self.head_dim ** 0.5),
                                        pred_mode=mode, orthogonal_matrix=orthogonal_matrix)
        self.core.anal = bool(anal) and bool(top_k)
        return self

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        if self.core is None:
            raise RuntimeError("MXSelfAttention.set_config(...) must be called first")
        B, N, C = hidden_states.shape
        q = self.to_q(hidden_states).view(B, N, self.num_heads, self.head_dim).transpose(1, 2)
        k = self.to_k(hidden_states).view(B, N, self.num_heads, self.head_dim).transpose(1, 2)
        v = self.to_v(hidden_states).view(B, N, self.num_heads, self.head_dim).transpose(1, 2)
        x = self.to_out(self.core(q, k, v, dense=self.current_timestep in self.exclude_timesteps))
        self.current_timestep += 1
        return x


class MXCrossAttention(MXSelfAttention):
    """PixArt-alpha cross-attention shim - mirrors workloads/PixArt/models/MX_transformer_block.py:720-859:
    keys / values come from the text encoder states (S tokens), and the additive attention_mask
    (B,1,S) is applied to the true and the predicted scores (:794-803, :821-822)."""

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        if self.core is None:
            raise RuntimeError("MXCrossAttention.set_config(...) must be called first")
        if encoder_hidden_states is None:
            raise ValueError("MXCrossAttention needs encoder_hidden_states")
        B, N, C = hidden_states.shape
        S = encoder_hidden_states.shape[1]
        q = self.to_q(hidden_states).view(B, N, self.num_heads, self.head_dim).transpose(1, 2)
        k = self.to_k(encoder_hidden_states).reshape(B, S, self.num_heads, self.head_dim).transpose(1, 2)
        v = self.to_v(encoder_hidden_states).reshape(B, S, self.num_heads, self.head_dim).transpose(1, 2)
        if attention_mask is not None and attention_mask.dtype == torch.bool:
            raise NotImplementedError("boolean masks (-inf bias) are not on the path; pass the additive fp32 mask")
        bias = None if attention_mask is None else attention_mask.to(torch.float32).reshape(B, S)
        excluded = self.current_timestep in self.exclude_timesteps
        x = self.to_out(self.core(q, k, v, key_bias=bias, pred_mode="exact" if excluded else None))
        self.current_timestep += 1
        return x


class MxLinear(nn.Linear):
    """Drop-in for the reference's mx.Linear (microxscaling/mx/linear.py:227-320) on the MXINT8
    configuration of the workloads: same constructor (in_features, out_features, bias, mx_specs),
    forward = ops.mx_linear.  The MX-quantized weight operand is cached and rebuilt when the weight
    tensor changes (inference: once)."""

    def __init__(self, in_features, out_features, bias=True, mx_specs=None, name=None):
        super().__init__(in_features, out_features, bias)
        resolve_linear_specs(mx_specs)          # also w_elem_format / round_weight / round (mx/linear.py:36-75)
        self.mx_specs = mx_specs
        self.name = name
        self._w_op = None
        self._w_key = None
        # a loaded checkpoint writes the weight through .data (no version bump): drop the cached operand
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module.invalidate())

    def invalidate(self):
        """Drop the cached MX-quantized weight operand (call after writing the weight through ``.data``)."""
        self._w_op = None
        self._w_key = None

    def forward(self, inputs):
        if torch.is_grad_enabled() and (self.weight.requires_grad or inputs.requires_grad):
            raise RuntimeError("MxLinear is the inference forward of mx.Linear (the reference's backward is out of scope): "
                               "call it under torch.no_grad() / inference_mode, or freeze the parameters")
        key = (self.weight.data_ptr(), self.weight._version, str(self.weight.device))
        if self._w_op is None or self._w_key != key:
            self._w_op = ops.mx_linear_prepare_weight(self.weight.detach(), self.mx_specs)
            self._w_key = key
        b = None if self.bias is None else self.bias.detach()
        return ops.mx_linear(inputs, self._w_op, b, self.mx_specs, out_features=self.out_features)

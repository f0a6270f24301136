"""The six prompt learners of TuneVLSeg, drop-in for ``src.models.core_models.coop.context_learner``.

These are tiny torch modules (a few thousand to a few hundred thousand parameters): their arithmetic stays in
torch so that they remain ordinary ``nn.Module``s for Lightning checkpointing, ``get_optim_groups`` name matching
and AdamW (SURVEY.md section 8b "Ownership").  Parameter names and ``state_dict`` layouts match the reference so
its checkpoints load; what differs is the interface the towers consume: ``visual_stack`` / ``textual_deep_stack``
hand the B200 tower kernels one (depth, n, D) table instead of being called back layer by layer.

Reference: base_unimodal_learner.py:17-99, coop_context_learner.py:15-181, base_projector_learner.py:10-139,
cocoop_context_learner.py:7-77, maple_context_learner.py:7-20, vpt_context_learner.py:15-64,
base_visual_learner.py:12-23, base_shared_learner.py:5-11, shared_attn_learner.py:9-104,
shared_separate_learner.py:11-98.
"""
from __future__ import annotations

import copy
import itertools
from abc import ABC, abstractmethod
from collections.abc import Iterable

import os

import torch
from torch import nn


class BaseUnimodalLearner(nn.Module, ABC):
    MIN_PROMPT_DEPTH = 1

    def __init__(self, *, max_network_depth: int, prompt_depth: int = 1, num_context: int | None = None,
                 context_dim: int | None = None, context_initializer=None, tokenizer=None, embedding_layer=None,
                 vector_std: float = 0.02, **kwargs) -> None:
        self.verify_prompt_depth(prompt_depth=prompt_depth, max_network_depth=max_network_depth)
        vectors = self.get_context_vectors(num_context=num_context, context_dim=context_dim,
                                           context_initializer=context_initializer, tokenizer=tokenizer,
                                           embedding_layer=embedding_layer, prompt_depth=prompt_depth,
                                           vector_std=vector_std)
        if vectors.ndim != 3:
            raise ValueError("The number of dimensions of `context_vectors` must be 3")
        if vectors.shape[0] != prompt_depth:
            raise ValueError("The number of rows of `context_vectors` must be `prompt_depth`")
        super().__init__()
        self.prompt_depth = prompt_depth
        self.num_context, self.context_dim = int(vectors.shape[1]), int(vectors.shape[2])
        self.context_vectors = nn.Parameter(vectors)

    @classmethod
    def verify_prompt_depth(cls, prompt_depth: int, max_network_depth: int) -> None:
        if prompt_depth < cls.MIN_PROMPT_DEPTH:
            raise ValueError(f"{prompt_depth=} must be at least {cls.MIN_PROMPT_DEPTH=}")
        if prompt_depth > max_network_depth:
            raise ValueError(f"{prompt_depth=} must be at most {max_network_depth=} for the used network.")

    @staticmethod
    def init_random_context_vectors(shape, std: float = 0.02) -> torch.Tensor:
        return nn.init.normal_(torch.empty(tuple(shape)), std=std)

    @abstractmethod
    def get_context_vectors(self, num_context=None, context_dim=None, prompt_depth=1, context_initializer=None,
                            tokenizer=None, embedding_layer=None, vector_std=0.02) -> torch.Tensor: ...


class CoOpContextLearner(BaseUnimodalLearner):
    """Learnable text context inserted after BOS (depth 0) and re-written into rows 1..n of deeper layers."""

    def get_context_vectors(self, num_context=None, context_dim=None, prompt_depth=1, context_initializer=None,
                            tokenizer=None, embedding_layer=None, vector_std=0.02) -> torch.Tensor:
        if context_initializer is None:
            if num_context is None or context_dim is None:
                raise ValueError("`num_context` and `context_dim` must be specified if `context_initializer` is None")
            return self.init_random_context_vectors((prompt_depth, num_context, context_dim), std=vector_std)
        if tokenizer is None or embedding_layer is None:
            raise ValueError("If `context_initializer` is not None, `tokenizer` and `embedding_layer` must be specified")
        text = context_initializer if isinstance(context_initializer, str) else context_initializer[:prompt_depth]
        seeded = self.get_context_vectors_from_initializer(text, embedding_layer, tokenizer)
        missing = prompt_depth - seeded.shape[0]
        if missing == 0:
            return seeded
        return torch.cat((seeded, self.init_random_context_vectors((missing, *seeded.shape[1:]), std=vector_std)))

    @staticmethod
    def get_context_vectors_from_initializer(context_initializer, embedding_layer, tokenizer) -> torch.Tensor:
        ids = tokenizer(context_initializer, return_tensors="pt", return_attention_mask=False, truncation=True,
                        add_special_tokens=False).input_ids
        with torch.no_grad():
            return embedding_layer(ids)

    # ---- mask helpers (coop_context_learner.py:82-114) -------------------------------------------------------
    def _update_mask_for_context(self, mask, constructor, max_length=None):
        lead = getattr(torch, constructor)(mask.shape[0], self.num_context, dtype=mask.dtype, device=mask.device)
        return torch.cat((lead, mask), dim=1)[:, :max_length]

    def update_attention_mask_for_context(self, attention_mask, max_length=None):
        return self._update_mask_for_context(attention_mask, "ones", max_length)

    def update_pad_mask_for_context(self, pad_mask, max_length=None):
        return self._update_mask_for_context(pad_mask, "zeros", max_length)

    # ---- contexts ----------------------------------------------------------------------------------------------
    def get_textual_context(self, in_context=None, image_features=None, index: int = 0) -> torch.Tensor:
        return self.context_vectors[index]

    def mutate_text_hidden_states(self, hidden_states, index: int, image_features=None):
        hidden_states[:, 1: self.num_context + 1] = self.get_textual_context(image_features=image_features, index=index)
        return hidden_states

    def textual_deep_stack(self, n_layers: int, image_features=None):
        """Contexts written after blocks 1 .. min(prompt_depth, n_layers) - 1, stacked on dim 0 (or None)."""
        last = min(self.prompt_depth, n_layers)
        if last <= 1:
            return None
        return torch.stack([self.get_textual_context(image_features=image_features, index=i) for i in range(1, last)])

    def forward(self, *, input_embeddings, max_length=None, image_features=None, context_vectors=None, index: int = 0):
        """[BOS, ctx x n, middle, last]; the middle is cut so that the total is min(L + n, max_length)."""
        n, L = self.num_context, input_embeddings.size(1)
        stop = -1 if max_length is None else min(max_length - n, L) - 1
        if context_vectors is None:
            context_vectors = self.get_textual_context(image_features=image_features, index=index)
        if context_vectors.ndim == 2:
            context_vectors = context_vectors.expand(input_embeddings.size(0), -1, -1)
        return torch.cat((input_embeddings[:, :1], context_vectors, input_embeddings[:, 1:stop], input_embeddings[:, -1:]), dim=1)


class BaseProjectorLearner(CoOpContextLearner):
    def __init__(self, *, proj_in_dim, proj_out_dim, prompt_depth: int = 1, use_unified_projection: bool = True,
                 intermediate_dim=None, use_proj_norm: bool = False, use_lora_proj: bool = False,
                 use_final_bias: bool = True, **kwargs) -> None:
        if use_lora_proj and intermediate_dim is not None and not isinstance(intermediate_dim, int):
            raise ValueError("Lora projection is only available for a single layer.")
        super().__init__(prompt_depth=prompt_depth, **kwargs)
        make = self.get_lora_projection if (use_lora_proj and intermediate_dim is not None) else self.get_mlp_projection
        kw = dict(in_dim=self.context_dim if proj_in_dim is None else proj_in_dim,
                  out_dim=self.context_dim if proj_out_dim is None else proj_out_dim,
                  intermediate_dim=intermediate_dim, use_final_norm=use_proj_norm, use_final_bias=use_final_bias)
        self.projection_layers = _layer_list(lambda: make(**kw), prompt_depth, use_unified_projection)

    def get_transformed_context(self, in_context=None, index: int = 0) -> torch.Tensor:
        return self.projection_layers[index](self.context_vectors[index] if in_context is None else in_context)

    @staticmethod
    def get_lora_projection(in_dim, out_dim, intermediate_dim, use_final_norm, use_final_bias: bool = True) -> nn.Sequential:
        layers = nn.Sequential(nn.Linear(in_dim, min(out_dim, intermediate_dim), bias=False))
        if intermediate_dim <= out_dim:
            layers.append(nn.Linear(intermediate_dim, out_dim, bias=(not use_final_norm) and use_final_bias))
        if use_final_norm:
            layers.append(nn.LayerNorm(out_dim, bias=use_final_bias))
        return layers

    @staticmethod
    def get_mlp_projection(in_dim, out_dim, intermediate_dim, use_final_norm, use_final_bias: bool = True):
        if intermediate_dim is None:
            return nn.Linear(in_dim, out_dim)
        dims = (intermediate_dim,) if isinstance(intermediate_dim, int) else tuple(intermediate_dim)
        layers = nn.Sequential(nn.Linear(in_dim, dims[0]), nn.ReLU(inplace=True))
        for i, o in itertools.pairwise(dims):
            layers.extend((nn.Linear(i, o), nn.ReLU(inplace=True)))
        for layer in layers:      # hidden layers: Kaiming-normal (base_projector_learner.py:118-122)
            if isinstance(layer, nn.Linear):
                nn.init.kaiming_normal_(layer.weight.data, nonlinearity="relu")
        layers.append(nn.Linear(dims[-1], out_dim, bias=(not use_final_norm) and use_final_bias))
        if use_final_norm:
            layers.append(nn.LayerNorm(out_dim, bias=use_final_bias))
        return layers


_STACK_DIRECT = os.environ.get("TVS_STACK_DIRECT", "1") != "0"      # A/B switch


class _StackParams(torch.autograd.Function):
    """``torch.stack`` of per-depth parameters whose backward adds the stacked gradient straight into the parameters'
    preallocated ``.grad`` buffers with ONE multi-tensor launch.

    With ``FusedAdamW`` every ``.grad`` is a view of the flat gradient buffer, so autograd's AccumulateGrad turns the
    backward of a plain ``torch.stack`` into one tiny ``add_`` launch per depth and parameter (~45 launches for MaPLe,
    all of them at the very end of the step's critical path, in front of the all-reduce).  Falls back to ordinary
    autograd behaviour when a gradient buffer is missing or a parameter is listed twice (unified projectors)."""

    @staticmethod
    def forward(ctx, *params):
        ctx.params = params
        return torch.stack([p.detach() for p in params])

    @staticmethod
    def backward(ctx, g):
        params = ctx.params
        slices = g.unbind(0)
        direct = (_STACK_DIRECT and len({id(p) for p in params}) == len(params)
                  and all(p.grad is not None and p.grad.dtype == g.dtype and p.grad.device == g.device for p in params))
        if direct:
            torch._foreach_add_([p.grad for p in params], list(slices))
            return (None,) * len(params)
        return tuple(slices)


def _batched_projection(layers, x: torch.Tensor):
    """Apply `len(layers)` structurally identical projectors to x[i] in a handful of batched launches.

    Every depth has its own tiny projector (n x 512 -> 64 -> 768 rows); evaluating them one by one costs ~15 kernel
    launches per depth in forward + backward, all on the step's critical path in front of the vision tower.  Stacking
    the per-depth parameters (they stay separate nn.Parameters: state_dict / optimizer groups are unchanged) turns
    them into two bmm's, one ReLU and one LayerNorm.  Returns None when the projectors are not of a batchable form
    (Linear | Linear-ReLU-Linear, optionally followed by LayerNorm).
    """
    seqs = [list(l) if isinstance(l, nn.Sequential) else [l] for l in layers]
    kinds = [tuple(type(m) for m in q) for q in seqs]
    if len(set(kinds)) != 1:
        return None
    kind = kinds[0]
    has_ln = kind and kind[-1] is nn.LayerNorm
    core = kind[:-1] if has_ln else kind
    if core not in ((nn.Linear,), (nn.Linear, nn.ReLU, nn.Linear)):
        return None

    def stk(ms, attr):
        ts = [getattr(m, attr) for m in ms]
        if any(t is None for t in ts):
            return None
        return _StackParams.apply(*ts) if all(t.requires_grad for t in ts) else torch.stack(ts)

    def lin(ms, inp):
        w, b = stk(ms, "weight"), stk(ms, "bias")
        out = torch.bmm(inp, w.transpose(1, 2))
        return out if b is None else out + b.unsqueeze(1)

    h = lin([q[0] for q in seqs], x)
    if len(core) == 3:
        h = lin([q[2] for q in seqs], torch.relu(h))
    if has_ln:
        lns = [q[-1] for q in seqs]
        if len({ln.eps for ln in lns}) != 1:
            return None
        h = torch.nn.functional.layer_norm(h, (h.shape[-1],), eps=lns[0].eps) * stk(lns, "weight").unsqueeze(1)
        beta = stk(lns, "bias")
        if beta is not None:
            h = h + beta.unsqueeze(1)
    return h


def _layer_list(factory, depth: int, unified: bool, clone_first: bool = False) -> nn.ModuleList:
    """One shared module repeated `depth` times (unified) or `depth` independent ones."""
    if unified:
        return nn.ModuleList((factory(),) * depth)
    if clone_first:
        first = factory()
        return nn.ModuleList(copy.deepcopy(first) for _ in range(depth))
    return nn.ModuleList(factory() for _ in range(depth))


class BaseVisualLearner(BaseUnimodalLearner):
    @abstractmethod
    def get_visual_context(self, in_context=None, index: int = 0) -> torch.Tensor: ...

    def mutate_image_hidden_states(self, hidden_states, index: int):
        hidden_states[:, -self.num_context:] = self.get_visual_context(index=index)
        return hidden_states

    def visual_stack(self, n_layers: int) -> torch.Tensor:
        """(prompt_depth, n, Dv) table: row 0 is concatenated, row idx is written after block idx < prompt_depth.
        Rows the tower never reads (idx >= n_layers + 1) are filled with zeros and get no gradient, exactly as the
        reference never evaluates them."""
        used = min(self.prompt_depth, n_layers + 1)
        rows = [self.get_visual_context(index=i) for i in range(used)]
        if used < self.prompt_depth:
            rows += [torch.zeros_like(rows[0])] * (self.prompt_depth - used)
        return torch.stack(rows)


class CoCoOpContextLearner(BaseProjectorLearner):
    def __init__(self, *, visual_dim: int, norm_image_features: bool = True, **kwargs) -> None:
        kwargs.update(proj_in_dim=visual_dim, proj_out_dim=None, use_final_bias=False)
        super().__init__(**kwargs)
        self.image_features_normalizer_or_identity = self._normalize_features if norm_image_features else nn.Identity()

    @staticmethod
    def _normalize_features(features, p="fro", dim: int = -1):
        return features / features.norm(p=p, dim=dim, keepdim=True)

    def get_textual_context(self, in_context=None, image_features=None, index: int = 0) -> torch.Tensor:
        if image_features is None:
            raise ValueError("`image_features` must be provided when `context_vectors` is None for CoCoOp")
        feats = self.image_features_normalizer_or_identity(image_features)
        shift = self.get_transformed_context(feats, index).unsqueeze(1)           # (B, 1, Dt)
        return shift + (self.context_vectors[index] if in_context is None else in_context)

    def forward(self, *, input_embeddings, max_length=None, image_features=None, context_vectors=None, index: int = 0):
        ctx = self.get_textual_context(in_context=context_vectors, image_features=image_features, index=index)
        return super().forward(input_embeddings=input_embeddings, max_length=max_length, context_vectors=ctx)


class MapleContextLearner(BaseProjectorLearner, BaseVisualLearner):
    def __init__(self, *, visual_dim: int, **kwargs) -> None:
        kwargs.update(proj_in_dim=None, proj_out_dim=visual_dim)
        super().__init__(**kwargs)

    def get_visual_context(self, *args, **kwargs) -> torch.Tensor:
        return self.get_transformed_context(*args, **kwargs)

    def visual_stack(self, n_layers: int) -> torch.Tensor:
        used = min(self.prompt_depth, n_layers + 1)
        out = _batched_projection(list(self.projection_layers)[:used], self.context_vectors[:used])
        if out is None:
            return super().visual_stack(n_layers)
        if used < self.prompt_depth:
            out = torch.cat((out, out.new_zeros((self.prompt_depth - used, *out.shape[1:]))))
        return out


class VPTContextLearner(BaseVisualLearner):
    def __init__(self, **kwargs) -> None:
        kwargs.update(context_initializer=None, tokenizer=None, embedding_layer=None)
        super().__init__(**kwargs)

    def get_context_vectors(self, num_context=None, context_dim=None, prompt_depth=1, context_initializer=None,
                            tokenizer=None, embedding_layer=None, vector_std=0.02) -> torch.Tensor:
        if num_context is None or context_dim is None:
            raise ValueError("`num_context` and `context_dim` must be specified for VPT")
        return self.init_random_context_vectors((prompt_depth, num_context, context_dim), std=vector_std)

    def get_visual_context(self, in_context=None, index: int = 0) -> torch.Tensor:
        return self.context_vectors[index]

    def forward(self, *, input_embeddings, max_length=None, image_features=None, context_vectors=None, index: int = 0):
        if context_vectors is None:
            context_vectors = self.context_vectors[index].expand(input_embeddings.size(0), -1, -1)
        return torch.cat((input_embeddings, context_vectors), dim=1)


class BaseSharedLearner(CoOpContextLearner, BaseVisualLearner):
    def __init__(self, **kwargs):
        kwargs.update(context_initializer=None, tokenizer=None, embedding_layer=None)
        super().__init__(**kwargs)


class SharedAttnLearner(BaseSharedLearner):
    """Shared (text|visual)-wide context pushed through one TransformerEncoderLayer per depth, then split.

    The reference computes the projection when the first branch asks for an index, returns that branch's half and
    parks the other half in a dict through the CPU (shared_attn_learner.py:43-92: a host sync per layer).  Here both
    halves stay on the device; the "compute once, consume once per branch" protocol is kept.
    """

    def __init__(self, *, textual_dim: int, visual_dim: int, unified_projector, prompt_depth: int = 1,
                 use_unified_projection: bool = True, **kwargs) -> None:
        if unified_projector is None:
            raise NotImplementedError("You need to provide a transformer encoder layer for the unified projection "
                                      "layer from the config.")
        kwargs["context_dim"] = textual_dim + visual_dim
        super().__init__(prompt_depth=prompt_depth, **kwargs)
        self.projection_layers = _layer_list(lambda: unified_projector(d_model=textual_dim + visual_dim), prompt_depth,
                                             use_unified_projection, clone_first=True)
        self._computed_textual_context_cache: dict[int, torch.Tensor] = {}
        self._computed_visual_context_cache: dict[int, torch.Tensor] = {}
        self.textual_dim, self.visual_dim = textual_dim, visual_dim

    def _get_combined_transformed_context(self, is_curr_branch_textual: bool, in_context=None, index: int = 0):
        mine = self._computed_textual_context_cache if is_curr_branch_textual else self._computed_visual_context_cache
        hit = mine.pop(index, None)
        if hit is not None:
            if hit.is_cuda:      # produced on the other branch's stream: keep the allocator from recycling it early
                hit.record_stream(torch.cuda.current_stream())
            return hit
        if in_context is None:
            in_context = self.context_vectors[index].unsqueeze(0)
        if in_context.ndim != 3:
            raise ValueError("The tensor needs to have 3 dimensions: (batch, context_len, hidden_dim)")
        both = self.projection_layers[index](in_context).squeeze(0)
        text_half, vis_half = both[:, : self.textual_dim], both[:, self.textual_dim:]
        if is_curr_branch_textual:
            self._computed_visual_context_cache[index] = vis_half
            return text_half
        self._computed_textual_context_cache[index] = text_half
        return vis_half

    def get_textual_context(self, image_features=None, *args, **kwargs) -> torch.Tensor:
        return self._get_combined_transformed_context(*args, is_curr_branch_textual=True, **kwargs)

    def get_visual_context(self, *args, **kwargs) -> torch.Tensor:
        return self._get_combined_transformed_context(*args, is_curr_branch_textual=False, **kwargs)


class SharedSeparateLearner(BaseSharedLearner):
    def __init__(self, *, textual_dim: int, visual_dim: int, shared_dim: int = 64, prompt_depth: int = 1,
                 use_unified_projection: bool = True, intermediate_dim: int | Iterable[int] | None = None,
                 use_proj_norm: bool = False, use_lora_proj: bool = False, **kwargs) -> None:
        if use_lora_proj and intermediate_dim is not None and not isinstance(intermediate_dim, int):
            raise ValueError("Lora projection is only available for a single layer.")
        kwargs["context_dim"] = shared_dim
        super().__init__(prompt_depth=prompt_depth, **kwargs)
        make = (BaseProjectorLearner.get_lora_projection if (use_lora_proj and intermediate_dim is not None)
                else BaseProjectorLearner.get_mlp_projection)
        kw = dict(in_dim=shared_dim, intermediate_dim=intermediate_dim, use_final_norm=use_proj_norm)
        self.textual_projection_layers = self.get_projection_layers(make(out_dim=textual_dim, **kw), prompt_depth,
                                                                    use_unified_projection)
        self.visual_projection_layers = self.get_projection_layers(make(out_dim=visual_dim, **kw), prompt_depth,
                                                                   use_unified_projection)

    @staticmethod
    def get_projection_layers(single_layer: nn.Module, prompt_depth: int, use_unified_projection) -> nn.ModuleList:
        if use_unified_projection:
            return nn.ModuleList((single_layer,) * prompt_depth)
        return nn.ModuleList(copy.deepcopy(single_layer) for _ in range(prompt_depth))

    def get_textual_context(self, in_context=None, image_features=None, index: int = 0) -> torch.Tensor:
        return self.textual_projection_layers[index](self.context_vectors[index] if in_context is None else in_context)

    def get_visual_context(self, in_context=None, index: int = 0) -> torch.Tensor:
        return self.visual_projection_layers[index](self.context_vectors[index] if in_context is None else in_context)

    def visual_stack(self, n_layers: int) -> torch.Tensor:
        used = min(self.prompt_depth, n_layers + 1)
        out = _batched_projection(list(self.visual_projection_layers)[:used], self.context_vectors[:used])
        if out is None:
            return super().visual_stack(n_layers)
        if used < self.prompt_depth:
            out = torch.cat((out, out.new_zeros((self.prompt_depth - used, *out.shape[1:]))))
        return out

    def textual_deep_stack(self, n_layers: int, image_features=None):
        last = min(self.prompt_depth, n_layers)
        if last <= 1:
            return None
        out = _batched_projection(list(self.textual_projection_layers)[1:last], self.context_vectors[1:last])
        return out if out is not None else super().textual_deep_stack(n_layers, image_features)

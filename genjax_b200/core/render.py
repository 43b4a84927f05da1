"""``render_html``: a jax-free stand-in for the reference's treescope rendering of pytrees
(src/genjax/_src/core/pytree.py:220-224 ``Pytree.render_html``): nested ``<details>`` blocks for choice maps, traces,
particle collections and filter results; tensors are summarised by dtype / shape / a few values.  Returns a string."""

from __future__ import annotations

import html

import torch


def _tensor(t: torch.Tensor) -> str:
    flat = t.detach().reshape(-1)
    head = ", ".join(f"{v:.6g}" for v in flat[:6].cpu().tolist())
    more = ", ..." if flat.numel() > 6 else ""
    stats = ""
    if flat.numel() > 1 and t.dtype.is_floating_point:
        f = flat.double()
        stats = f" mean={f.mean().item():.4g} sd={f.std().item():.4g}"
    return f"<code>{str(t.dtype).replace('torch.', '')}{list(t.shape)}</code> [{head}{more}]{stats}"


def _node(label: str, body: str, open_: bool = True) -> str:
    return f"<details{' open' if open_ else ''}><summary>{html.escape(label)}</summary><div style='margin-left:1.2em'>{body}</div></details>"


def _render(obj, depth: int = 0) -> str:
    from .choice_map import ChoiceMap

    if isinstance(obj, torch.Tensor):
        return _tensor(obj)
    if isinstance(obj, ChoiceMap):
        rows = []
        for addr, v in obj.leaves():
            a = "/".join(str(x) for x in (addr if isinstance(addr, tuple) else (addr,)))
            rows.append(f"<div><b>{html.escape(a)}</b>: {_render(getattr(v, 'value', v), depth + 1)}</div>")
        return _node(f"ChoiceMap ({len(rows)} addresses)", "".join(rows) or "<i>empty</i>", depth < 2)
    if hasattr(obj, "get_choices") and hasattr(obj, "get_score"):  # a trace
        body = f"<div><b>score</b>: {_render(obj.get_score(), depth + 1)}</div>"
        try:
            body += f"<div><b>retval</b>: {_render(obj.get_retval(), depth + 1)}</div>"
        except Exception:
            pass
        body += _render(obj.get_choices(), depth + 1)
        name = getattr(getattr(obj, "get_gen_fn", lambda: None)(), "__name__", type(obj).__name__)
        return _node(f"Trace of {name}", body, depth < 2)
    if hasattr(obj, "get_particles") and hasattr(obj, "get_log_weights"):  # ParticleCollection
        lw = obj.get_log_weights()
        body = (f"<div><b>log_weights</b>: {_tensor(lw)}</div><div><b>log marginal likelihood estimate</b>: "
                f"{float(obj.get_log_marginal_likelihood_estimate()):.6g}</div><div><b>ESS</b>: {float(obj.effective_sample_size()):.6g}"
                f" of {lw.numel()}</div>" + _render(obj.get_particles(), depth + 1))
        return _node("ParticleCollection", body, depth < 2)
    if hasattr(obj, "log_increments") and hasattr(obj, "state"):  # PFResult
        body = (f"<div><b>log marginal likelihood</b>: {float(obj.log_marginal_likelihood):.6g}</div>"
                f"<div><b>log increments</b>: {_tensor(obj.log_increments)}</div>"
                + "".join(f"<div><b>state[{i}]</b>: {_tensor(s)}</div>" for i, s in enumerate(obj.state)))
        if obj.ancestors is not None:
            body += f"<div><b>ESS per step</b>: {_tensor(obj.ess)}</div><div><b>ancestors</b>: {_tensor(obj.ancestors)}</div>"
        return _node("PFResult", body, depth < 2)
    if isinstance(obj, (tuple, list)):
        return _node(f"{type(obj).__name__}[{len(obj)}]", "".join(f"<div>{_render(o, depth + 1)}</div>" for o in obj), depth < 2)
    if isinstance(obj, dict):
        return _node(f"dict[{len(obj)}]", "".join(f"<div><b>{html.escape(str(k))}</b>: {_render(v, depth + 1)}</div>" for k, v in obj.items()), depth < 2)
    return f"<code>{html.escape(repr(obj))}</code>"


def render_html(obj) -> str:
    """HTML rendering of a choice map, trace, particle collection, filter result or a (nested) container of them."""
    return "<div style='font-family:monospace;font-size:13px'>" + _render(obj) + "</div>"

"""torch.optim.Adam restated on plain tensors (ORACLE, test infra).

Follows torch/optim/adam.py `_single_tensor_adam` (the CPU default: foreach=False, amsgrad=False,
weight_decay=0, maximize=False) op for op -- the reference builds every optimiser with
`self.opt = torch.optim.Adam` (rltoolkit/rl.py:62) and default betas/eps:

    exp_avg.lerp_(grad, 1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    denom = (exp_avg_sq.sqrt() / sqrt(1 - beta2**t)).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-(lr / (1 - beta1**t)))
"""
import torch

BETA1, BETA2, EPS = 0.9, 0.999, 1e-8


def adam_step(param: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor,
              step: int, lr: float) -> None:
    """In-place Adam step; `step` is the 1-based step count AFTER the increment."""
    m.lerp_(grad, 1 - BETA1)
    v.mul_(BETA2).addcmul_(grad, grad, value=1 - BETA2)
    bias_correction1 = 1 - BETA1 ** step
    bias_correction2 = 1 - BETA2 ** step
    step_size = lr / bias_correction1
    bias_correction2_sqrt = bias_correction2 ** 0.5
    denom = (v.sqrt() / bias_correction2_sqrt).add_(EPS)
    param.addcdiv_(m, denom, value=-step_size)


def adam_step_net(state: dict, net: str, grads: dict, lr: float) -> None:
    """Adam over every tensor of `net` (keys '<net>.<name>'), moments at '<key>#m' / '#v',
    shared step counter at '<net>#step' (torch keeps one per tensor; they move in lockstep)."""
    step = int(state.get(net + "#step", 0)) + 1
    state[net + "#step"] = step
    for name, g in grads.items():
        key = net + "." + name
        if key + "#m" not in state:
            state[key + "#m"] = torch.zeros_like(state[key])
            state[key + "#v"] = torch.zeros_like(state[key])
        adam_step(state[key], g.reshape(state[key].shape), state[key + "#m"], state[key + "#v"],
                  step, lr)

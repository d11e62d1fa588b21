"""FusedAdam: drop-in for the reference optimizer (torch_bindings/adam.py:6-36).

Subclass of torch.optim.Adam with one tensor per param group, state keys `step` / `exp_avg` /
`exp_avg_sq` (Model.py mutates them through NeRFICG's adam_utils), `.step()` without closure that
launches one dense Adam kernel per group through the C ABI (`fgs_adam_step`)."""
import torch

from .. import _C


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr, eps) -> None:
        super().__init__(params=params, lr=lr, eps=eps)

    def _state_for(self, param: torch.Tensor) -> dict:
        state = self.state[param]
        if len(state) == 0:
            state['step'] = 0
            state['exp_avg'] = torch.zeros_like(param)
            state['exp_avg_sq'] = torch.zeros_like(param)
        return state

    @torch.no_grad()
    def step(self) -> None:
        for group in self.param_groups:
            self.step_group(group)

    @torch.no_grad()
    def step_group(self, group: dict) -> None:
        """The update of ONE param group (what `step` does for each of them).  The view-parallel path calls it
        tensor by tensor, as soon as that tensor's gradient all-reduce has completed (fgs_b200/parallel.py)."""
        assert len(group['params']) == 1, 'more than one tensor in group'
        param = group['params'][0]
        if param.grad is None or param.numel() == 0:
            return
        state = self._state_for(param)
        state['step'] += 1
        _C.adam_step(param.grad, param, state['exp_avg'], state['exp_avg_sq'], state['step'], group['lr'],
                     *group['betas'], group['eps'])

    def group_of(self, param: torch.Tensor) -> dict:
        for group in self.param_groups:
            if any(p is param for p in group['params']):
                return group
        raise KeyError('tensor is not managed by this optimizer')

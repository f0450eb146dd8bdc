"""The two metrics the training loop calls every iteration (train.py:24 of the
reference passes [accuracy, dice]; definitions at utils/metrics.py:31-47,114-130)."""
import numpy as np
import torch


def accuracy(P, G):
    if torch.is_tensor(P) and torch.is_tensor(G):
        return (P == G).float().mean().item()
    return (np.asarray(P) == np.asarray(G)).mean()


def dice(S, G, epsilon=1e-7):
    if torch.is_tensor(S) and torch.is_tensor(G):
        S = S.unsqueeze(0) if S.dim() == 2 else S
        G = G.unsqueeze(0) if G.dim() == 2 else G
        S, G = S.float(), G.float()
        score = 2 * (G * S).sum(dim=(1, 2)) / (G.sum(dim=(1, 2)) + S.sum(dim=(1, 2)) + epsilon)
        return score.mean().item()
    S, G = np.asarray(S, dtype=np.float64), np.asarray(G, dtype=np.float64)
    if S.ndim == 2:
        S, G = S[None], G[None]
    score = 2 * (G * S).sum(axis=(1, 2)) / (G.sum(axis=(1, 2)) + S.sum(axis=(1, 2)) + epsilon)
    return float(score.mean())


# Device-side variants used by the trainer loop: same arithmetic, but the result stays
# a 0-dim tensor so the iteration can read all of its scalars in one transfer.
def _accuracy_t(P, G):
    return (P == G).float().mean()


def _dice_t(S, G, epsilon=1e-7):
    S = S.unsqueeze(0) if S.dim() == 2 else S
    G = G.unsqueeze(0) if G.dim() == 2 else G
    S, G = S.float(), G.float()
    return (2 * (G * S).sum(dim=(1, 2)) / (G.sum(dim=(1, 2)) + S.sum(dim=(1, 2)) + epsilon)).mean()


accuracy.deferred = _accuracy_t
dice.deferred = _dice_t

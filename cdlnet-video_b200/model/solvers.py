"""Constructor-time spectral normalisation and train-time projection
(reference model/solvers.py:3-28).  Boundary only: they run on the modules' stock
nn.Conv attributes, once, and are not kernel targets."""
import torch


def power_method(A, b, num_iter=1000, tol=1e-6, verbose=True):
    """Largest eigenvalue of the operator A by power iteration from b.
    Returns (eigenvalue, eigenvector, tolerance_reached)."""
    prev = None
    reached = False
    eig = torch.zeros(())
    for it in range(num_iter):
        b = A(b)
        b = b / torch.norm(b)
        eig = torch.sum(b * A(b))
        delta = abs(float(eig) - prev) if prev is not None else abs(float(eig))
        if verbose:
            print(f"i:{it:3d} \t |e_new - e_old|:{delta:2.2e}")
        if delta < tol:
            reached = True
            break
        prev = float(eig)
    if verbose:
        print("tolerance reached!", it)
        print(f"L = {float(eig):.3e}")
    return float(eig), b, reached


def uball_project(W, dim=(2, 3)):
    """Project each filter onto the unit l2 ball."""
    norm = torch.norm(W, dim=dim, keepdim=True)
    return W * torch.clamp(1.0 / norm, max=1)

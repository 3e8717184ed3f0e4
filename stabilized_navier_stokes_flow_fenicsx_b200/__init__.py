"""B200-native stabilized Navier-Stokes residual/Jacobian assembly + CSR SpMV (hot path only)."""

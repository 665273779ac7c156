"""trixib200: B200-native TreeMesh DGSEM rhs! behind the TrixiCUDA.jl API (host mirror in Python)."""

"""Stand-in for the third-party `pointnet2_ops` package, backed by the CPU oracle (test infrastructure)."""

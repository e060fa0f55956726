"""mpl_toolkits stub (test tooling only)."""

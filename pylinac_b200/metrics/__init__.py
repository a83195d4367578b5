"""Metric plug-ins (mirror of pylinac/metrics): profile metrics computed on the GPU-backed profile classes."""

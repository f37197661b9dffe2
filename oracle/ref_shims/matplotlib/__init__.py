"""Stub of matplotlib: the reference env only asks for a colormap callable at load time."""

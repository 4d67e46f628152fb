"""Import-only stub (real turtle needs tkinter)."""
def update(*a, **k): pass

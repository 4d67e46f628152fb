"""Import-only stub (tests/golden/gen_golden.py): the reference imports `ic` for debug prints."""
def ic(*a, **k):
    return a[0] if a else None

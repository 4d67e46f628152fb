"""Import-only stub for golden generation; never executed on the CoAlign path."""

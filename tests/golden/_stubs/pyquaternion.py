"""Import-only stub."""
class Quaternion: pass

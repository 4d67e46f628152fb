"""Import-only stub."""
class Polygon: pass
class Point: pass
class MultiPoint: pass

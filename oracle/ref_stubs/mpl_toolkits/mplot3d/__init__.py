class Axes3D:  # imported by name only
    pass

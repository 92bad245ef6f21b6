"""placeholder, filled in below"""
def write_scene_header(scene, flat):
    return "(scene header generation not yet implemented)"

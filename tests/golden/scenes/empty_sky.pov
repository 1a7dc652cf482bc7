// no objects at all: every primary ray ends in ComputeSky (sky_sphere over the background)
#version 3.7;
global_settings { assumed_gamma 1 }
camera { location <0, 1, -5> look_at <0, 1.5, 0> angle 60 right x*16/9 }
light_source { <5, 10, -5> rgb 1 }
background { rgb <0.1, 0.1, 0.3> }
sky_sphere { pigment { gradient y color_map { [0 rgb <0.9, 0.8, 0.6>] [0.5 rgbt <0.3, 0.5, 0.9, 0.4>] [1 rgb <0.05, 0.1, 0.4>] } scale 2 translate -1 } }

// cam_spherical: non-pinhole camera types of TracePixel::CreateCameraRay (tracepixel.cpp:394-674)
#version 3.7;
camera { spherical location <0, 2.0, -1.0> look_at <0, 1.5, 4> angle 300 120 right x*16/9 }
global_settings { assumed_gamma 1 max_trace_level 3 }
light_source { <10, 15, -12> rgb <1, 1, 1> }
background { rgb <0.1, 0.15, 0.3> }
plane { y, 0 pigment { checker rgb <0.9, 0.9, 0.9>, rgb <0.3, 0.3, 0.35> } finish { ambient 0.1 diffuse 0.7 } }
sphere { <-2.5, 1.0, 3.0>, 1.0 pigment { rgb <0.9, 0.3, 0.2> } finish { ambient 0.1 diffuse 0.7 phong 0.5 } }
sphere { <2.8, 1.4, 2.0>, 1.4 pigment { rgbf <0.8, 1.0, 0.8, 0.7> } finish { ambient 0.05 diffuse 0.3 specular 0.4 reflection 0.1 } interior { ior 1.3 } }
box { <-1.0, 0, 5.0>, <1.0, 2.5, 6.5> pigment { rgb <0.3, 0.4, 0.9> } finish { ambient 0.1 diffuse 0.7 } rotate y*20 }
cylinder { <0, 0, -4>, <0, 3, -4>, 0.6 pigment { rgb <0.9, 0.8, 0.2> } finish { ambient 0.1 diffuse 0.7 } }
sphere { <0, 6, 0>, 1.0 pigment { rgb <0.8, 0.8, 0.8> } finish { ambient 0.1 diffuse 0.6 reflection 0.3 } }

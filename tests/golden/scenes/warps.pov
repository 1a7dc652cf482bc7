// the point-mapping warps (warp.cpp:124-545): black_hole (plain, inverse, repeating), repeat (with flip and offset), cubic,
// cylindrical, spherical, toroidal and planar warps in front of patterns and image-like gradients
#version 3.7;
global_settings { assumed_gamma 1.0 }
camera { location <0, 0, -11> look_at <0, 0, 0> angle 48 }
light_source { <-10, 12, -20> rgb 1 }
#declare CM = color_map { [0 rgb <0.05, 0.05, 0.3>] [0.3 rgb <0.2, 0.6, 0.9>] [0.6 rgb <1, 0.9, 0.3>] [1 rgb <0.9, 0.2, 0.1>] }
#macro Tile(X, Y, P)
  box { <-1, -1, 0>, <1, 1, 0.1> pigment { P } finish { ambient 0.3 diffuse 0.7 } translate <X, Y, 0> }
#end
Tile(-3.3,  2.2, pigment { checker rgb 0.1, rgb 0.9 scale 0.25 warp { black_hole <0.2, 0.1, 0>, 0.8 strength 1.5 falloff 2 } })
Tile(-1.1,  2.2, pigment { checker rgb <0.9, 0.3, 0.2>, rgb 0.9 scale 0.2 warp { black_hole <0, 0, 0>, 0.6 inverse strength 0.8 falloff 3 repeat <0.9, 0.9, 0> } })
Tile( 1.1,  2.2, pigment { gradient x color_map { CM } scale 0.5 warp { repeat 0.7 * x offset <0, 0.2, 0> flip x } })
Tile( 3.3,  2.2, pigment { bozo color_map { CM } scale 0.3 warp { repeat 0.5 * y flip <1, 1, 0> } warp { turbulence 0.2 } })
Tile(-3.3,  0.0, pigment { gradient x color_map { CM } frequency 4 warp { cylindrical orientation y dist_exp 1 } translate <0.1, 0, -0.5> })
Tile(-1.1,  0.0, pigment { checker rgb 0.1, rgb <0.4, 0.9, 0.5> scale 0.1 warp { spherical orientation z dist_exp 0.5 } translate <0.2, -0.1, 0.8> })
Tile( 1.1,  0.0, pigment { gradient y color_map { CM } frequency 3 warp { toroidal orientation z dist_exp 0 major_radius 0.7 } rotate 90 * x translate <0, 0, 0.3> })
Tile( 3.3,  0.0, pigment { checker rgb <0.2, 0.3, 0.8>, rgb 0.95 scale 0.2 warp { planar <0.3, 1, 0.2>, 0.4 } })
Tile(-3.3, -2.2, pigment { gradient x color_map { CM } frequency 6 warp { cubic } translate <0.1, 0.2, -0.7> })
Tile(-1.1, -2.2, pigment { marble color_map { CM } scale 0.4 warp { cylindrical } warp { repeat 0.8 * x } translate <0, 0, -0.4> })
sphere { <1.6, -2.2, 0>, 1 pigment { checker rgb 0.1, rgb <0.9, 0.7, 0.2> scale <0.05, 0.1, 1> warp { spherical } } finish { phong 0.5 } }
torus { 0.75, 0.3 pigment { gradient x color_map { CM } frequency 8 warp { toroidal major_radius 0.75 } } rotate -70 * x translate <3.6, -2.2, 0> }

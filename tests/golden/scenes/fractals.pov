// the FractalPattern family (mandel / julia / magnet with exponents 2-4, every exterior and interior colouring type that is cheap to
// show) and spiral1 / spiral2 with and without classic turbulence (pattern.cpp:6895-7751, 8396-8517, 8990-9056)
#version 3.7;
global_settings { assumed_gamma 1.0 }
camera { location <0, 0, -11> look_at <0, 0, 0> angle 48 }
light_source { <-10, 12, -20> rgb 1 }
#declare CM = color_map { [0 rgb <0.05, 0.05, 0.3>] [0.15 rgb <0.2, 0.6, 0.9>] [0.4 rgb <1, 0.9, 0.3>] [0.7 rgb <0.9, 0.2, 0.1>] [1 rgb <0.05, 0, 0>] }
#macro Tile(X, Y, P)
  box { <-1, -1, 0>, <1, 1, 0.1> pigment { P } finish { ambient 0.3 diffuse 0.7 } translate <X, Y, 0> }
#end
Tile(-3.3,  2.2, pigment { mandel 40 color_map { CM } scale 0.55 translate <0.35, 0, 0> })
Tile(-1.1,  2.2, pigment { mandel 30 exponent 3 exterior 5, 0.2 interior 1, 2 color_map { CM } scale 0.6 })
Tile( 1.1,  2.2, pigment { mandel 25 exponent 4 exterior 6, 0.15 interior 4, 3 color_map { CM } scale 0.6 })
Tile( 3.3,  2.2, pigment { julia <0.353, 0.288>, 30 interior 1, 1 color_map { CM } scale 0.65 })
Tile(-3.3,  0.0, pigment { julia <0.4, 0.2>, 20 exponent 3 exterior 7, 5 color_map { CM } scale 0.7 })
Tile(-1.1,  0.0, pigment { julia <0.5, 0.1>, 20 exponent 4 exterior 8, 4 interior 6, 0.3 color_map { CM } scale 0.7 })
Tile( 1.1,  0.0, pigment { magnet 1 mandel 30 interior 2, 0.5 color_map { CM } scale 0.3 translate <-0.5, 0, 0> })
Tile( 3.3,  0.0, pigment { magnet 2 mandel 20 exterior 2, 0.01 color_map { CM } scale 0.35 translate <-0.5, 0, 0> })
Tile(-3.3, -2.2, pigment { magnet 1 julia <0.3, 0.5>, 25 exterior 3, 0.02 interior 3, 1 color_map { CM } scale 0.4 })
Tile(-1.1, -2.2, pigment { magnet 2 julia <0.2, 0.6>, 20 interior 5, 2 color_map { CM } scale 0.4 })
Tile( 1.1, -2.2, pigment { spiral1 3 color_map { CM } scale 0.4 })
Tile( 3.3, -2.2, pigment { spiral2 5 turbulence 0.3 octaves 3 color_map { CM } frequency 2 sine_wave scale 0.5 rotate 20 * z })

// Regenerates the golden outputs from the REAL reference dependencies (see README.md).
// Not compiled in the build image (no Go toolchain there).
package main

import (
	"encoding/json"
	"fmt"
	"image"
	"image/color"
	"image/draw"
	"os"
	"path/filepath"

	xdraw "golang.org/x/image/draw"
)

type glyph struct {
	X0, Y0, X1, Y1, MpX, MpY, MaskW, MaskH int
	Mask                                   string
}
type kase struct {
	Name, Kind string
	W, H       int
	Planes     []string
	Ops        [][]interface{}
	Color      []int
	Glyphs     []glyph
}

func must(err error) {
	if err != nil {
		panic(err)
	}
}
func read(dir, name string) []byte { b, err := os.ReadFile(filepath.Join(dir, name)); must(err); return b }

func source(dir string, k kase) image.Image {
	r := image.Rect(0, 0, k.W, k.H)
	switch k.Kind {
	case "rgba", "rgba_premul":
		return &image.RGBA{Pix: read(dir, k.Planes[0]), Stride: 4 * k.W, Rect: r}
	case "nrgba":
		return &image.NRGBA{Pix: read(dir, k.Planes[0]), Stride: 4 * k.W, Rect: r}
	case "gray":
		return &image.Gray{Pix: read(dir, k.Planes[0]), Stride: k.W, Rect: r}
	}
	ratio := map[string]image.YCbCrSubsampleRatio{"ycbcr444": image.YCbCrSubsampleRatio444, "ycbcr422": image.YCbCrSubsampleRatio422,
		"ycbcr420": image.YCbCrSubsampleRatio420, "ycbcr440": image.YCbCrSubsampleRatio440}[k.Kind]
	m := image.NewYCbCr(r, ratio)
	copy(m.Y, read(dir, k.Planes[0]))
	copy(m.Cb, read(dir, k.Planes[1]))
	copy(m.Cr, read(dir, k.Planes[2]))
	return m
}

// operations/resize.go:121-125
func resizeImage(img image.Image, w, h int) *image.RGBA {
	dst := image.NewRGBA(image.Rect(0, 0, w, h))
	xdraw.BiLinear.Scale(dst, dst.Bounds(), img, img.Bounds(), xdraw.Over, nil)
	return dst
}

// operations/thumbnail.go:114-132
func cropAndResize(img image.Image, size int) *image.RGBA {
	b := img.Bounds()
	ow, oh := b.Dx(), b.Dy()
	var cx, cy, cs int
	if ow > oh {
		cs, cx, cy = oh, (ow-oh)/2, 0
	} else {
		cs, cx, cy = ow, 0, (oh-ow)/2
	}
	cropped := image.NewRGBA(image.Rect(0, 0, cs, cs))
	xdraw.BiLinear.Scale(cropped, cropped.Bounds(), img, image.Rect(cx, cy, cx+cs, cy+cs), xdraw.Over, nil)
	return resizeImage(cropped, size, size)
}

func main() {
	dir := os.Args[1]
	var cases []kase
	must(json.Unmarshal(read(dir, "index.json"), &cases))
	for _, k := range cases {
		img := source(dir, k)
		for _, op := range k.Ops {
			var out *image.RGBA
			var tag string
			switch op[0].(string) {
			case "resize":
				w, h := int(op[1].(float64)), int(op[2].(float64))
				out, tag = resizeImage(img, w, h), fmt.Sprintf("resize_%dx%d", w, h)
			case "thumb":
				s := int(op[1].(float64))
				out, tag = cropAndResize(img, s), fmt.Sprintf("thumb_%d", s)
			case "blend": // operations/watermark.go:91-92 + the per-rune DrawMask freetype issues
				out = image.NewRGBA(img.Bounds())
				draw.Draw(out, out.Bounds(), img, image.Point{}, draw.Src)
				src := image.NewUniform(color.RGBA{uint8(k.Color[0]), uint8(k.Color[1]), uint8(k.Color[2]), uint8(k.Color[3])})
				for _, g := range k.Glyphs {
					mask := &image.Alpha{Pix: read(dir, g.Mask), Stride: g.MaskW, Rect: image.Rect(0, 0, g.MaskW, g.MaskH)}
					draw.DrawMask(out, image.Rect(g.X0, g.Y0, g.X1, g.Y1), src, image.Point{}, mask, image.Pt(g.MpX, g.MpY), draw.Over)
				}
				tag = "blend"
			}
			must(os.WriteFile(filepath.Join(dir, k.Name+"."+tag+".out"), out.Pix, 0o644))
		}
	}
}

// image_codecs.cc -- ImageIO without FreeImage.
//
// The reference decodes and encodes images through a FreeImage.dll it loads at run time
// (raylib/render/image.cc:152-263, raylib/loader/dll_loader.h:23-35).  That DLL does not exist off Windows, so
// the formats the renderer's callers actually use are implemented here directly:
//   read : PNG (zlib inflate; 8/16-bit, grey / RGB / palette / alpha, non-interlaced), BMP (24/32-bit BI_RGB),
//          Radiance HDR (RGBE, flat or RLE), PPM/PGM (P2 P3 P5 P6), PFM, TGA, JPEG (jpeg_codec.cc)
//   write: BMP (24-bit), PNG (8-bit RGB), JPEG (baseline 4:2:0, quality 75), PPM (by ".ppm" extension)
// Conventions follow the reference loader: LDR texels become Pixel(r,g,b,a) = byte / 255 with row 0 at the TOP
// of the picture (image.cc:214-226), HDR texels are copied as floats with alpha 1 (image.cc:169-195).
// Host-side media I/O is outside the GPU hot path.
#include "render/image.h"
#include "core/logger.h"

#include <zlib.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace
{
	typedef std::vector<unsigned char> Bytes;

	bool ReadFile(const char* path, Bytes& out)
	{
		FILE* f = fopen(path, "rb");
		if (!f) return false;
		fseek(f, 0, SEEK_END);
		const long n = ftell(f);
		fseek(f, 0, SEEK_SET);
		out.resize(n > 0 ? (size_t)n : 0);
		const bool ok = n >= 0 && fread(out.data(), 1, out.size(), f) == out.size();
		fclose(f);
		return ok;
	}

	std::string LowerExtension(const char* path)
	{
		const char* dot = strrchr(path, '.');
		std::string e = dot ? dot + 1 : "";
		for (char& c : e) c = (char)tolower((unsigned char)c);
		return e;
	}

	uint32 BE32(const unsigned char* p) { return ((uint32)p[0] << 24) | ((uint32)p[1] << 16) | ((uint32)p[2] << 8) | p[3]; }
	uint32 LE32(const unsigned char* p) { return ((uint32)p[3] << 24) | ((uint32)p[2] << 16) | ((uint32)p[1] << 8) | p[0]; }
	uint32 LE16(const unsigned char* p) { return ((uint32)p[1] << 8) | p[0]; }

	Image2D* FromRGBA8(uint32 w, uint32 h, const Bytes& rgba)
	{
		Image2D* image = new Image2D(w, h);
		Pixel* dst = image->MutablePixels();
		for (size_t i = 0; i < (size_t)w * h; ++i)
			dst[i] = Pixel((uint8)rgba[4 * i], (uint8)rgba[4 * i + 1], (uint8)rgba[4 * i + 2], (uint8)rgba[4 * i + 3]);
		return image;
	}

	// ---- PNG ---------------------------------------------------------------------------------------
	int Paeth(int a, int b, int c)
	{
		const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
		return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
	}

	Image2D* LoadPNG(const Bytes& file)
	{
		static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
		if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) return nullptr;
		uint32 w = 0, h = 0, depth = 0, colorType = 0, interlace = 0;
		Bytes idat, palette, trns;
		size_t pos = 8;
		while (pos + 12 <= file.size())
		{
			const uint32 len = BE32(&file[pos]);
			const unsigned char* type = &file[pos + 4];
			const unsigned char* data = &file[pos + 8];
			if (pos + 12 + (size_t)len > file.size()) return nullptr;
			if (!memcmp(type, "IHDR", 4) && len >= 13) { w = BE32(data); h = BE32(data + 4); depth = data[8]; colorType = data[9]; interlace = data[12]; }
			else if (!memcmp(type, "PLTE", 4)) palette.assign(data, data + len);
			else if (!memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
			else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
			else if (!memcmp(type, "IEND", 4)) break;
			pos += 12 + (size_t)len;
		}
		// bit depths the format allows for each colour type (grey 1/2/4/8/16, palette 1/2/4/8, everything else 8/16); anything
		// else would divide by zero or shift by a negative amount below.  Dimensions are capped so that a hostile header
		// cannot ask for terabytes (the row buffer below is (stride + 1) * h bytes).
		const bool depthOk = colorType == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
		                   : colorType == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
		                   : (depth == 8 || depth == 16);
		if (w == 0 || h == 0 || w > 65536u || h > 65536u || interlace != 0 || !depthOk) return nullptr;
		const uint32 channels = colorType == 0 ? 1 : colorType == 2 ? 3 : colorType == 3 ? 1 : colorType == 4 ? 2 : colorType == 6 ? 4 : 0;
		if (channels == 0) return nullptr;
		const size_t bitsPerPixel = (size_t)channels * depth;
		const size_t stride = ((size_t)w * bitsPerPixel + 7) / 8, bpp = std::max<size_t>(1, bitsPerPixel / 8);
		if ((stride + 1) > (size_t)1 << 31 || (stride + 1) * (size_t)h > (size_t)1 << 33) return nullptr;
		Bytes raw((stride + 1) * h);
		uLongf rawLen = (uLongf)raw.size();
		if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) return nullptr;
		// undo the per-scanline filters in place
		Bytes prev(stride, 0);
		for (uint32 y = 0; y < h; ++y)
		{
			unsigned char* line = &raw[(stride + 1) * y];
			const int filter = line[0];
			unsigned char* cur = line + 1;
			for (size_t i = 0; i < stride; ++i)
			{
				const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
				int v = cur[i];
				switch (filter) { case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) / 2; break; case 4: v += Paeth(a, b, c); break; default: break; }
				cur[i] = (unsigned char)v;
			}
			memcpy(prev.data(), cur, stride);
		}
		Bytes rgba((size_t)w * h * 4);
		for (uint32 y = 0; y < h; ++y)
		{
			const unsigned char* cur = &raw[(stride + 1) * y + 1];
			for (uint32 x = 0; x < w; ++x)
			{
				unsigned char* o = &rgba[4 * ((size_t)y * w + x)];
				auto sample = [&](uint32 c) -> uint32 {
					if (depth == 8) return cur[(size_t)x * channels + c];
					if (depth == 16) return cur[2 * ((size_t)x * channels + c)];       // high byte
					const uint32 bit = x * depth, v = (cur[bit / 8] >> (8 - depth - (bit % 8))) & ((1u << depth) - 1u);
					return colorType == 3 ? v : v * 255u / ((1u << depth) - 1u);
				};
				if (colorType == 3)
				{
					const uint32 idx = sample(0);
					for (int c = 0; c < 3; ++c) o[c] = 3 * idx + c < palette.size() ? palette[3 * idx + c] : 0;
					o[3] = idx < trns.size() ? trns[idx] : 255;
				}
				else if (colorType == 0) { o[0] = o[1] = o[2] = (unsigned char)sample(0); o[3] = 255; }
				else if (colorType == 4) { o[0] = o[1] = o[2] = (unsigned char)sample(0); o[3] = (unsigned char)sample(1); }
				else { o[0] = (unsigned char)sample(0); o[1] = (unsigned char)sample(1); o[2] = (unsigned char)sample(2); o[3] = colorType == 6 ? (unsigned char)sample(3) : 255; }
			}
		}
		return FromRGBA8(w, h, rgba);
	}

	void PutBE32(Bytes& b, uint32 v) { b.push_back((unsigned char)(v >> 24)); b.push_back((unsigned char)(v >> 16)); b.push_back((unsigned char)(v >> 8)); b.push_back((unsigned char)v); }
	void PutChunk(Bytes& out, const char* type, const Bytes& data)
	{
		PutBE32(out, (uint32)data.size());
		const size_t start = out.size();
		out.insert(out.end(), type, type + 4);
		out.insert(out.end(), data.begin(), data.end());
		PutBE32(out, (uint32)crc32(0L, &out[start], (uInt)(out.size() - start)));
	}

	bool WritePNG(const Image2D* image, const char* path)
	{
		const uint32 w = image->GetWidth(), h = image->GetHeight();
		Bytes raw;
		raw.reserve(((size_t)w * 3 + 1) * h);
		for (uint32 y = 0; y < h; ++y)
		{
			raw.push_back(0);
			for (uint32 x = 0; x < w; ++x)
			{
				const uint32 argb = image->GetPixel((int32)x, (int32)y).ToUint32();
				raw.push_back((unsigned char)(argb >> 16)); raw.push_back((unsigned char)(argb >> 8)); raw.push_back((unsigned char)argb);
			}
		}
		Bytes z(compressBound((uLong)raw.size()));
		uLongf zLen = (uLongf)z.size();
		if (compress2(z.data(), &zLen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
		z.resize(zLen);
		Bytes out = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
		Bytes ihdr;
		PutBE32(ihdr, w); PutBE32(ihdr, h);
		const unsigned char tail[5] = { 8, 2, 0, 0, 0 };
		ihdr.insert(ihdr.end(), tail, tail + 5);
		PutChunk(out, "IHDR", ihdr);
		PutChunk(out, "IDAT", z);
		PutChunk(out, "IEND", Bytes());
		FILE* f = fopen(path, "wb");
		if (!f) return false;
		const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
		fclose(f);
		return ok;
	}

	// ---- BMP ---------------------------------------------------------------------------------------
	Image2D* LoadBMP(const Bytes& file)
	{
		if (file.size() < 54 || file[0] != 'B' || file[1] != 'M') return nullptr;
		const uint32 offset = LE32(&file[10]), headerSize = LE32(&file[14]);
		const int32 w = (int32)LE32(&file[18]), hRaw = (int32)LE32(&file[22]);
		const uint32 bits = LE16(&file[28]), compression = LE32(&file[30]);
		if (headerSize < 40 || w <= 0 || hRaw == 0 || (bits != 24 && bits != 32) || (compression != 0 && compression != 3)) return nullptr;
		const uint32 h = (uint32)abs(hRaw);
		const size_t stride = (((size_t)w * bits + 31) / 32) * 4;
		if (offset + stride * h > file.size()) return nullptr;
		Bytes rgba((size_t)w * h * 4);
		for (uint32 y = 0; y < h; ++y)
		{
			const unsigned char* src = &file[offset + stride * (hRaw > 0 ? (h - 1 - y) : y)];      // positive height = bottom-up
			for (int32 x = 0; x < w; ++x)
			{
				const unsigned char* p = src + (size_t)x * (bits / 8);
				unsigned char* o = &rgba[4 * ((size_t)y * w + x)];
				o[0] = p[2]; o[1] = p[1]; o[2] = p[0]; o[3] = bits == 32 ? p[3] : 255;
			}
		}
		return FromRGBA8((uint32)w, h, rgba);
	}

	bool WriteBMP(const Image2D* image, const char* path)
	{
		const uint32 w = image->GetWidth(), h = image->GetHeight();
		const uint32 stride = ((w * 3 + 3) / 4) * 4;
		Bytes out(54 + (size_t)stride * h, 0);
		auto put32 = [&](size_t at, uint32 v) { out[at] = (unsigned char)v; out[at + 1] = (unsigned char)(v >> 8); out[at + 2] = (unsigned char)(v >> 16); out[at + 3] = (unsigned char)(v >> 24); };
		out[0] = 'B'; out[1] = 'M';
		put32(2, (uint32)out.size()); put32(10, 54); put32(14, 40); put32(18, w); put32(22, h);
		out[26] = 1; out[28] = 24; put32(34, stride * h);
		for (uint32 y = 0; y < h; ++y)
			for (uint32 x = 0; x < w; ++x)
			{
				const uint32 argb = image->GetPixel((int32)x, (int32)(h - 1 - y)).ToUint32();
				unsigned char* p = &out[54 + (size_t)stride * y + 3 * x];
				p[0] = (unsigned char)argb; p[1] = (unsigned char)(argb >> 8); p[2] = (unsigned char)(argb >> 16);
			}
		FILE* f = fopen(path, "wb");
		if (!f) return false;
		const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
		fclose(f);
		return ok;
	}

	// ---- Radiance HDR ----------------------------------------------------------------------------------
	Image2D* LoadHDR(const Bytes& file)
	{
		if (file.size() < 11 || (memcmp(file.data(), "#?RADIANCE", 10) != 0 && memcmp(file.data(), "#?RGBE", 6) != 0)) return nullptr;
		size_t pos = 0;
		auto line = [&]() { std::string s; while (pos < file.size() && file[pos] != '\n') s.push_back((char)file[pos++]); if (pos < file.size()) ++pos; return s; };
		for (;;) { if (pos >= file.size()) return nullptr; if (line().empty()) break; }
		const std::string res = line();
		int w = 0, h = 0;
		if (sscanf(res.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) return nullptr;
		Image2D* image = new Image2D((uint32)w, (uint32)h);
		Pixel* dst = image->MutablePixels();
		Bytes scan((size_t)w * 4);
		for (int y = 0; y < h; ++y)
		{
			if (pos + 4 > file.size()) { delete image; return nullptr; }
			if (w >= 8 && w < 32768 && file[pos] == 2 && file[pos + 1] == 2 && (((int)file[pos + 2] << 8) | file[pos + 3]) == w)
			{
				pos += 4;
				for (int c = 0; c < 4; ++c)
					for (int x = 0; x < w;)
					{
						if (pos >= file.size()) { delete image; return nullptr; }
						int count = file[pos++];
						if (count > 128) { count -= 128; if (pos >= file.size() || x + count > w) { delete image; return nullptr; } const unsigned char v = file[pos++]; while (count--) scan[4 * (size_t)x++ + c] = v; }
						else { if (count == 0 || pos + count > file.size() || x + count > w) { delete image; return nullptr; } while (count--) scan[4 * (size_t)x++ + c] = file[pos++]; }
					}
			}
			else
			{
				if (pos + (size_t)w * 4 > file.size()) { delete image; return nullptr; }
				memcpy(scan.data(), &file[pos], (size_t)w * 4);
				pos += (size_t)w * 4;
			}
			for (int x = 0; x < w; ++x)
			{
				const unsigned char* p = &scan[4 * (size_t)x];
				const float f = p[3] ? std::ldexp(1.0f, (int)p[3] - (128 + 8)) : 0.0f;
				dst[(size_t)y * w + x] = Pixel(p[0] * f, p[1] * f, p[2] * f, 1.0f);
			}
		}
		return image;
	}

	// ---- PNM / PFM -------------------------------------------------------------------------------------
	Image2D* LoadPNM(const Bytes& file)
	{
		if (file.size() < 3 || file[0] != 'P') return nullptr;
		const char kind = (char)file[1];
		size_t pos = 2;
		auto token = [&]() {
			std::string s;
			for (;;)
			{
				while (pos < file.size() && isspace(file[pos])) ++pos;
				if (pos < file.size() && file[pos] == '#') { while (pos < file.size() && file[pos] != '\n') ++pos; continue; }
				break;
			}
			while (pos < file.size() && !isspace(file[pos])) s.push_back((char)file[pos++]);
			return s;
		};
		if (kind == 'f' || kind == 'F')
		{
			const int w = atoi(token().c_str()), h = atoi(token().c_str());
			const float scale = (float)atof(token().c_str());
			++pos;
			const int ch = kind == 'F' ? 3 : 1;
			if (w <= 0 || h <= 0 || pos + (size_t)w * h * ch * 4 > file.size()) return nullptr;
			Image2D* image = new Image2D((uint32)w, (uint32)h);
			for (int y = 0; y < h; ++y)
				for (int x = 0; x < w; ++x)
				{
					float v[3];
					for (int c = 0; c < ch; ++c)
					{
						unsigned char b[4];
						memcpy(b, &file[pos + 4 * (((size_t)(h - 1 - y) * w + x) * ch + c)], 4);      // PFM rows run bottom-up
						if (scale > 0.0f) std::swap(b[0], b[3]), std::swap(b[1], b[2]);                 // big-endian file
						memcpy(&v[c], b, 4);
					}
					image->MutablePixels()[(size_t)y * w + x] = ch == 3 ? Pixel(v[0], v[1], v[2], 1.0f) : Pixel(v[0], v[0], v[0], 1.0f);
				}
			return image;
		}
		if (kind != '2' && kind != '3' && kind != '5' && kind != '6') return nullptr;
		const int w = atoi(token().c_str()), h = atoi(token().c_str()), maxv = atoi(token().c_str());
		if (w <= 0 || h <= 0 || maxv <= 0 || maxv > 65535) return nullptr;
		const int ch = (kind == '3' || kind == '6') ? 3 : 1;
		const bool binary = kind == '5' || kind == '6';
		if (binary) ++pos;
		Bytes rgba((size_t)w * h * 4);
		for (size_t i = 0; i < (size_t)w * h; ++i)
		{
			int v[3] = { 0, 0, 0 };
			for (int c = 0; c < ch; ++c)
			{
				if (binary)
				{
					if (pos >= file.size()) return nullptr;
					v[c] = file[pos++];
					if (maxv > 255) { if (pos >= file.size()) return nullptr; v[c] = (v[c] << 8) | file[pos++]; }
				}
				else v[c] = atoi(token().c_str());
				v[c] = v[c] * 255 / maxv;
			}
			if (ch == 1) v[1] = v[2] = v[0];
			rgba[4 * i] = (unsigned char)v[0]; rgba[4 * i + 1] = (unsigned char)v[1]; rgba[4 * i + 2] = (unsigned char)v[2]; rgba[4 * i + 3] = 255;
		}
		return FromRGBA8((uint32)w, (uint32)h, rgba);
	}

	bool WritePPM(const Image2D* image, const char* path)
	{
		FILE* f = fopen(path, "wb");
		if (!f) return false;
		fprintf(f, "P6\n%u %u\n255\n", image->GetWidth(), image->GetHeight());
		for (uint32 y = 0; y < image->GetHeight(); ++y)
			for (uint32 x = 0; x < image->GetWidth(); ++x)
			{
				const uint32 argb = image->GetPixel((int32)x, (int32)y).ToUint32();
				const unsigned char rgb[3] = { (unsigned char)(argb >> 16), (unsigned char)(argb >> 8), (unsigned char)argb };
				fwrite(rgb, 1, 3, f);
			}
		fclose(f);
		return true;
	}
}

namespace
{
	// ---- TGA (types 2 / 3 / 10 / 11: true colour and grey, raw or run-length encoded; 8, 24 and 32 bits) -----------------
	// Sponza-class OBJ scenes ship their textures as .tga; FreeImage decodes them for the reference (render/image.cc:152-230).
	Image2D* LoadTGA(const Bytes& f)
	{
		if (f.size() < 18) return nullptr;
		const uint32 idLength = f[0], colorMapType = f[1], type = f[2];
		const uint32 w = f[12] | (f[13] << 8), h = f[14] | (f[15] << 8), bits = f[16], descriptor = f[17];
		if (colorMapType != 0 || w == 0 || h == 0) return nullptr;
		const bool rle = type == 10 || type == 11, grey = type == 3 || type == 11;
		if (!(type == 2 || type == 3 || rle)) return nullptr;
		if (!((grey && bits == 8) || (!grey && (bits == 24 || bits == 32)))) return nullptr;
		const size_t bpp = bits / 8, count = (size_t)w * h;
		size_t pos = 18 + idLength;
		Bytes px(count * bpp);
		if (!rle)
		{
			if (pos + px.size() > f.size()) return nullptr;
			memcpy(px.data(), &f[pos], px.size());
		}
		else
		{
			size_t out = 0;
			while (out < count)
			{
				if (pos >= f.size()) return nullptr;
				const uint32 head = f[pos++], run = (head & 0x7f) + 1;
				if (out + run > count) return nullptr;
				if (head & 0x80)
				{
					if (pos + bpp > f.size()) return nullptr;
					for (uint32 k = 0; k < run; ++k) memcpy(&px[(out + k) * bpp], &f[pos], bpp);
					pos += bpp;
				}
				else
				{
					if (pos + run * bpp > f.size()) return nullptr;
					memcpy(&px[out * bpp], &f[pos], run * bpp);
					pos += run * bpp;
				}
				out += run;
			}
		}
		const bool topDown = (descriptor & 0x20) != 0, rightToLeft = (descriptor & 0x10) != 0;
		Image2D* image = new Image2D(w, h, 0xff000000);
		for (uint32 y = 0; y < h; ++y)
			for (uint32 x = 0; x < w; ++x)
			{
				const size_t sy = topDown ? y : h - 1 - y, sx = rightToLeft ? w - 1 - x : x;
				const unsigned char* p = &px[(sy * w + sx) * bpp];
				const float k = 1.0f / 255.0f;
				if (grey) image->SetPixel((int32)x, (int32)y, Pixel(p[0] * k, p[0] * k, p[0] * k, 1.0f));
				else image->SetPixel((int32)x, (int32)y, Pixel(p[2] * k, p[1] * k, p[0] * k, bpp == 4 ? p[3] * k : 1.0f));     // stored B, G, R(, A)
			}
		return image;
	}
}

Image2D* RtLoadJPEG(const std::vector<unsigned char>& file, const char** why);      // jpeg_codec.cc
bool RtWriteJPEG(const Image2D* image, const char* path, int quality);

namespace ImageIO
{
	// Reference: render/image.cc:152-230 (FreeImage::GetFIFFromFilename + Load + ConvertTo32Bits / ConvertToRGBAF).
	Image2D* LoadImage2DFromFile(const char* filepath)
	{
		if (filepath == nullptr) return nullptr;
		Bytes file;
		if (!ReadFile(filepath, file)) { LOG("ImageIO: cannot read '%s'", filepath); return nullptr; }
		const std::string ext = LowerExtension(filepath);
		Image2D* image = nullptr;
		if (ext == "png") image = LoadPNG(file);
		else if (ext == "bmp") image = LoadBMP(file);
		else if (ext == "hdr" || ext == "pic") image = LoadHDR(file);
		else if (ext == "ppm" || ext == "pgm" || ext == "pnm" || ext == "pfm") image = LoadPNM(file);
		else if (ext == "tga") image = LoadTGA(file);
		else if (ext == "jpg" || ext == "jpeg" || ext == "jpe" || ext == "jfif")
		{
			const char* why = nullptr;
			image = RtLoadJPEG(file, &why);
			if (!image) LOG("ImageIO: JPEG decoder: %s", why ? why : "failed");
		}
		else
		{
			// unknown extension: go by the file's magic number
			image = LoadPNG(file);
			if (!image && file.size() > 2 && file[0] == 0xFF && file[1] == 0xD8) image = RtLoadJPEG(file, nullptr);
			if (!image) image = LoadBMP(file);
			if (!image) image = LoadHDR(file);
			if (!image) image = LoadPNM(file);
		}
		if (!image)
		{
			// loud on purpose: a texture that fails to load turns into a constant-colour material, i.e. a wrong picture
			LOG("ImageIO: '%s' is not a PNG / BMP / TGA / JPEG / HDR / PNM file this build can decode", filepath);
			fprintf(stderr, "raylib-b200: cannot decode image '%s' (built-in codecs: PNG, BMP, TGA, Huffman JPEG, Radiance HDR, PNM/PFM)\n", filepath);
		}
		return image;
	}

	// Reference: render/image.cc:232-263 (24-bit RGB through FreeImage::Save).
	bool WriteImage2DToDisk(Image2D* image, const char* filepath, EImageFileType fileType)
	{
		if (!image || !filepath || image->GetWidth() == 0 || image->GetHeight() == 0) return false;
		if (LowerExtension(filepath) == "ppm") return WritePPM(image, filepath);
		switch (fileType)
		{
		case RAYLIB_IMAGEFILETYPE_Bitmap: return WriteBMP(image, filepath);
		case RAYLIB_IMAGEFILETYPE_Png: return WritePNG(image, filepath);
		case RAYLIB_IMAGEFILETYPE_Jpg: return RtWriteJPEG(image, filepath, 75);      // FreeImage::Save(..., 0): JPEG_DEFAULT = quality 75, 4:2:0
		default:
			LOG("ImageIO: cannot write '%s': unknown file type %d", filepath, (int)fileType);
			return false;
		}
	}
}

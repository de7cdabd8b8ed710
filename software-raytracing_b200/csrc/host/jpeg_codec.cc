// jpeg_codec.cc -- JPEG without FreeImage: a decoder for the files OBJ scenes ship as textures and an encoder for the
// pictures the reference's client writes (src/main.cc:478-510 saves every result as .jpg through
// Raylib_WriteImageToDisk -> ImageIO::WriteImage2DToDisk, render/image.cc:232-263, FreeImage::Save with flags 0 =
// quality 75, 4:2:0).
//
// Decoder: baseline, extended-sequential and progressive Huffman JPEG (SOF0 / SOF1 / SOF2), 8-bit samples, one (grey)
// or three (YCbCr, or RGB when an Adobe marker / component ids say so) components, any sampling factors, restart
// intervals, 8- and 16-bit quantization tables.  The sample pipeline restates libjpeg's defaults (the library behind
// FreeImage): the accurate integer inverse DCT ("islow": 13-bit constants, two passes), the 16-bit fixed-point
// YCbCr -> RGB tables, and for subsampled chroma the "fancy" triangle filters of libjpeg 6b / libjpeg-turbo (2:1
// horizontal, 2x2, 1x2).  tests/test_cpu_host.py compares the output with libjpeg-turbo's byte by byte; files without
// chroma subsampling decode to the same bytes under every libjpeg version.
// Not handled (refused, never guessed): arithmetic coding, lossless, 12-bit, CMYK / YCCK.
//
// Encoder: baseline 4:2:0 with the Annex K quantization tables scaled to quality 75 and the Annex K Huffman tables.
// Host-side media I/O is outside the GPU hot path.
#include "render/image.h"
#include "core/logger.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

namespace
{
	typedef std::vector<unsigned char> Bytes;

	const int kZigzag[64] = {
		0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
		35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };

	// ---- Huffman tables ------------------------------------------------------------------------------
	struct HuffTable
	{
		bool present = false;
		unsigned char bits[17] = {};        // bits[l] = number of codes of length l
		unsigned char vals[256] = {};
		int maxCode[18], valPtr[17], minCode[17];
		unsigned short look[512];           // 9-bit look-ahead: (length << 8) | symbol, 0 = longer code

		bool Prepare()
		{
			int code = 0, k = 0;
			memset(look, 0, sizeof(look));
			for (int l = 1; l <= 16; ++l)
			{
				valPtr[l] = k; minCode[l] = code;
				for (int i = 0; i < bits[l]; ++i, ++k, ++code)
				{
					if (k >= 256) return false;
					if (l <= 9)
						for (int fill = 0; fill < (1 << (9 - l)); ++fill) look[(code << (9 - l)) | fill] = (unsigned short)((l << 8) | vals[k]);
				}
				maxCode[l] = bits[l] ? code - 1 : -1;
				if (code > (1 << l)) return false;          // more codes than the length can hold
				code <<= 1;
			}
			maxCode[17] = 0x7FFFFFFF;
			present = true;
			return true;
		}
	};

	// ---- entropy-coded segment reader -----------------------------------------------------------------
	struct BitReader
	{
		const unsigned char* p; const unsigned char* end;
		unsigned int acc = 0; int count = 0;
		bool atMarker = false;

		void Fill()
		{
			while (count <= 24)
			{
				unsigned int b = 0;
				if (!atMarker && p < end)
				{
					b = *p;
					if (b == 0xFF)
					{
						if (p + 1 < end && p[1] == 0x00) p += 2;          // stuffed zero
						else if (p + 1 < end && p[1] == 0xFF) { ++p; continue; }      // fill bytes before a marker
						else { atMarker = true; b = 0; }
					}
					else ++p;
				}
				acc |= b << (24 - count);
				count += 8;
			}
		}
		int Get(int n)
		{
			if (n == 0) return 0;
			if (count < n) Fill();
			const int v = (int)(acc >> (32 - n));
			acc <<= n; count -= n;
			return v;
		}
		int Peek9() { if (count < 9) Fill(); return (int)(acc >> 23); }
		void Skip(int n) { acc <<= n; count -= n; }
		int Decode(const HuffTable& h)
		{
			const unsigned short fast = h.look[Peek9()];
			if (fast) { Skip(fast >> 8); return fast & 0xFF; }
			if (count < 16) Fill();
			int code = (int)(acc >> 23), l = 9;
			while (l < 16 && code > h.maxCode[l]) { ++l; code = (int)(acc >> (32 - l)); }
			if (l > 16 || code > h.maxCode[l] || h.maxCode[l] < 0) { Skip(16); return 0; }      // corrupt data: keep going, the picture shows it
			Skip(l);
			return h.vals[(h.valPtr[l] + code - h.minCode[l]) & 0xFF];
		}
		static int Extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }
		int Receive(int s) { return s ? Extend(Get(s), s) : 0; }
		// byte-align and step over the RSTn marker that should be next
		void Restart()
		{
			acc = 0; count = 0; atMarker = false;
			while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7))
			{
				if (p[0] == 0xFF && p[1] != 0x00 && p[1] != 0xFF) return;      // some other marker: leave it to the segment parser
				++p;
			}
			if (p + 1 < end) p += 2;
		}
	};

	struct Component
	{
		int id = 0, h = 1, v = 1, tq = 0;
		int td = 0, ta = 0;                  // tables of the current scan
		int blocksW = 0, blocksH = 0;        // padded to whole MCUs (coefficient storage)
		int scanW = 0, scanH = 0;            // block grid of a scan that holds this component alone
		int sampleW = 0, sampleH = 0;        // libjpeg's downsampled_width / _height
		int pred = 0;
		std::vector<short> coef;             // 64 per block, natural order
		std::vector<unsigned char> plane;    // blocksW*8 x blocksH*8 samples
	};

	// libjpeg's jidctint.c (accurate integer inverse DCT), restated: quantized input, samples out.  Intermediates are
	// 64-bit: a valid stream never leaves 32 bits (libjpeg relies on that), a damaged one must not overflow either.
	typedef long long I64;
	inline void IdctOddEven(const I64* d, I64* sum, I64* diff)
	{
		// even part
		const I64 z1 = (d[2] + d[6]) * 4433;
		const I64 e2 = z1 + d[6] * -15137, e3 = z1 + d[2] * 6270;
		const I64 e0 = (d[0] + d[4]) * 8192, e1 = (d[0] - d[4]) * 8192;
		const I64 t10 = e0 + e3, t13 = e0 - e3, t11 = e1 + e2, t12 = e1 - e2;
		// odd part
		I64 t0 = d[7], t1 = d[5], t2 = d[3], t3 = d[1];
		I64 y1 = t0 + t3, y2 = t1 + t2, y3 = t0 + t2, y4 = t1 + t3;
		const I64 y5 = (y3 + y4) * 9633;
		t0 *= 2446; t1 *= 16819; t2 *= 25172; t3 *= 12299;
		y1 *= -7373; y2 *= -20995; y3 *= -16069; y4 *= -3196;
		y3 += y5; y4 += y5;
		t0 += y1 + y3; t1 += y2 + y4; t2 += y2 + y3; t3 += y1 + y4;
		sum[0] = t10 + t3; diff[0] = t10 - t3;      // outputs 0 / 7
		sum[1] = t11 + t2; diff[1] = t11 - t2;      // 1 / 6
		sum[2] = t12 + t1; diff[2] = t12 - t1;      // 2 / 5
		sum[3] = t13 + t0; diff[3] = t13 - t0;      // 3 / 4
	}
	void InverseDCT(const short* in, const unsigned short* quant, unsigned char* out, int stride)
	{
		const int C = 13, P = 2;
		I64 ws[64];
		for (int c = 0; c < 8; ++c)
		{
			I64 d[8], sum[4], diff[4];
			for (int k = 0; k < 8; ++k) d[k] = (I64)in[8 * k + c] * (I64)quant[8 * k + c];
			IdctOddEven(d, sum, diff);
			const I64 r = 1 << (C - P - 1); const int s = C - P;
			for (int k = 0; k < 4; ++k) { ws[8 * k + c] = (sum[k] + r) >> s; ws[8 * (7 - k) + c] = (diff[k] + r) >> s; }
		}
		for (int row = 0; row < 8; ++row)
		{
			I64 sum[4], diff[4];
			IdctOddEven(ws + 8 * row, sum, diff);
			const int s = C + P + 3; const I64 r = (I64)1 << (s - 1);
			for (int k = 0; k < 4; ++k)
			{
				const I64 a = ((sum[k] + r) >> s) + 128, b = ((diff[k] + r) >> s) + 128;
				out[row * stride + k] = (unsigned char)std::min<I64>(255, std::max<I64>(0, a));
				out[row * stride + 7 - k] = (unsigned char)std::min<I64>(255, std::max<I64>(0, b));
			}
		}
	}

	struct Decoder
	{
		const Bytes& file;
		size_t pos = 2;
		int width = 0, height = 0, numComps = 0, hMax = 1, vMax = 1, mcusX = 0, mcusY = 0;
		bool progressive = false, sawFrame = false;
		int adobeTransform = -1;
		bool jfif = false;
		int restartInterval = 0;
		unsigned short quant[4][64];
		bool quantPresent[4] = { false, false, false, false };
		HuffTable dc[4], ac[4];
		Component comps[3];
		const char* error = nullptr;

		explicit Decoder(const Bytes& f) : file(f) {}

		bool Fail(const char* why) { if (!error) error = why; return false; }

		bool ParseTables(int marker, size_t at, size_t len)
		{
			const unsigned char* d = &file[at];
			if (marker == 0xDB)
			{
				size_t i = 0;
				while (i < len)
				{
					const int pq = d[i] >> 4, tq = d[i] & 15; ++i;
					if (tq > 3 || pq > 1 || i + (pq ? 128 : 64) > len) return Fail("bad quantization table");
					for (int k = 0; k < 64; ++k)
					{
						quant[tq][kZigzag[k]] = pq ? (unsigned short)((d[i] << 8) | d[i + 1]) : d[i];
						i += pq ? 2 : 1;
					}
					quantPresent[tq] = true;
				}
			}
			else if (marker == 0xC4)
			{
				size_t i = 0;
				while (i < len)
				{
					if (i + 17 > len) return Fail("bad Huffman table");
					const int tc = d[i] >> 4, th = d[i] & 15; ++i;
					if (tc > 1 || th > 3) return Fail("bad Huffman table id");
					HuffTable& h = tc ? ac[th] : dc[th];
					int total = 0;
					h.bits[0] = 0;
					for (int l = 1; l <= 16; ++l) { h.bits[l] = d[i++]; total += h.bits[l]; }
					if (total > 256 || i + total > len) return Fail("bad Huffman table size");
					memset(h.vals, 0, sizeof(h.vals));
					memcpy(h.vals, d + i, total); i += total;
					if (!h.Prepare()) return Fail("inconsistent Huffman table");
				}
			}
			else if (marker == 0xDD) { if (len >= 2) restartInterval = (d[0] << 8) | d[1]; }
			else if (marker == 0xEE) { if (len >= 12 && memcmp(d, "Adobe", 5) == 0) adobeTransform = d[11]; }
			else if (marker == 0xE0) { if (len >= 5 && memcmp(d, "JFIF", 5) == 0) jfif = true; }
			return true;
		}

		bool ParseFrame(int marker, size_t at, size_t len)
		{
			const unsigned char* d = &file[at];
			if (sawFrame) return Fail("more than one frame");
			if (len < 6 || d[0] != 8) return Fail("only 8-bit samples are supported");
			height = (d[1] << 8) | d[2]; width = (d[3] << 8) | d[4]; numComps = d[5];
			if (width <= 0 || height <= 0 || width > 32768 || height > 32768) return Fail("unsupported picture size");
			if (!(numComps == 1 || numComps == 3) || len < (size_t)(6 + 3 * numComps)) return Fail("only grey and three-component pictures are supported");
			progressive = marker == 0xC2;
			for (int c = 0; c < numComps; ++c)
			{
				Component& k = comps[c];
				k.id = d[6 + 3 * c]; k.h = d[7 + 3 * c] >> 4; k.v = d[7 + 3 * c] & 15; k.tq = d[8 + 3 * c] & 3;
				if (k.h < 1 || k.h > 4 || k.v < 1 || k.v > 4) return Fail("bad sampling factors");
				hMax = std::max(hMax, k.h); vMax = std::max(vMax, k.v);
			}
			if (numComps == 1) { comps[0].h = comps[0].v = 1; hMax = vMax = 1; }      // a lone component is never interleaved
			mcusX = (width + 8 * hMax - 1) / (8 * hMax); mcusY = (height + 8 * vMax - 1) / (8 * vMax);
			for (int c = 0; c < numComps; ++c)
			{
				Component& k = comps[c];
				k.blocksW = mcusX * k.h; k.blocksH = mcusY * k.v;
				k.sampleW = (width * k.h + hMax - 1) / hMax; k.sampleH = (height * k.v + vMax - 1) / vMax;
				k.scanW = (k.sampleW + 7) / 8; k.scanH = (k.sampleH + 7) / 8;
				k.coef.assign((size_t)k.blocksW * k.blocksH * 64, 0);
			}
			sawFrame = true;
			return true;
		}

		// one 8x8 block of one scan
		void DecodeBlock(BitReader& br, Component& k, short* b, int ss, int se, int ah, int al, int& eobRun)
		{
			if (!progressive)
			{
				const int t = br.Decode(dc[k.td]);
				k.pred = std::min(32767, std::max(-32768, k.pred + br.Receive(t & 15)));      // (the clamp only ever acts on damaged data)
				b[0] = (short)k.pred;
				for (int i = 1; i < 64; ++i)
				{
					const int rs = br.Decode(ac[k.ta]), r = rs >> 4, s = rs & 15;
					if (s == 0) { if (r != 15) break; i += 15; continue; }
					i += r;
					if (i > 63) break;
					b[kZigzag[i]] = (short)br.Receive(s);
				}
				return;
			}
			if (ss == 0)
			{
				if (ah == 0)
				{
					const int t = br.Decode(dc[k.td]);
					k.pred = std::min(32767, std::max(-32768, k.pred + br.Receive(t & 15)));
					b[0] = (short)(k.pred * (1 << al));
				}
				else if (br.Get(1)) b[0] |= (short)(1 << al);
				return;
			}
			if (ah == 0)
			{
				if (eobRun > 0) { --eobRun; return; }
				for (int i = ss; i <= se; ++i)
				{
					const int rs = br.Decode(ac[k.ta]), r = rs >> 4, s = rs & 15;
					if (s == 0)
					{
						if (r < 15) { eobRun = (1 << r) - 1; if (r) eobRun += br.Get(r); break; }
						i += 15; continue;
					}
					i += r;
					if (i > 63) break;
					b[kZigzag[i]] = (short)(br.Receive(s) * (1 << al));
				}
				return;
			}
			// successive-approximation refinement of the AC band
			const int p1 = 1 << al, m1 = -(1 << al);
			int i = ss;
			if (eobRun == 0)
			{
				for (; i <= se; ++i)
				{
					const int rs = br.Decode(ac[k.ta]);
					int r = rs >> 4, s = rs & 15;
					if (s) s = br.Get(1) ? p1 : m1;
					else if (r != 15) { eobRun = 1 << r; if (r) eobRun += br.Get(r); break; }
					do
					{
						short& c = b[kZigzag[i]];
						if (c != 0) { if (br.Get(1) && (c & p1) == 0) c = (short)(c + (c >= 0 ? p1 : m1)); }
						else if (--r < 0) break;
						++i;
					} while (i <= se);
					if (s && i <= 63) b[kZigzag[i]] = (short)s;
				}
			}
			if (eobRun > 0)
			{
				for (; i <= se; ++i)
				{
					short& c = b[kZigzag[i]];
					if (c != 0 && br.Get(1) && (c & p1) == 0) c = (short)(c + (c >= 0 ? p1 : m1));
				}
				--eobRun;
			}
		}

		// returns the position of the marker that ends the scan
		bool DecodeScan(size_t at, size_t len, size_t& next)
		{
			const unsigned char* d = &file[at];
			if (!sawFrame || len < 1) return Fail("scan before frame");
			const int ns = d[0];
			if (ns < 1 || ns > numComps || len < (size_t)(4 + 2 * ns)) return Fail("bad scan header");
			Component* scan[3];
			for (int i = 0; i < ns; ++i)
			{
				scan[i] = nullptr;
				for (int c = 0; c < numComps; ++c) if (comps[c].id == d[1 + 2 * i]) scan[i] = &comps[c];
				if (!scan[i]) return Fail("scan names an unknown component");
				scan[i]->td = d[2 + 2 * i] >> 4; scan[i]->ta = d[2 + 2 * i] & 15;
				if (scan[i]->td > 3 || scan[i]->ta > 3) return Fail("bad table selector");
			}
			int ss = d[1 + 2 * ns], se = d[2 + 2 * ns];
			const int ah = d[3 + 2 * ns] >> 4, al = d[3 + 2 * ns] & 15;
			if (!progressive) { ss = 0; se = 63; }
			if (ss > se || se > 63 || (progressive && ss > 0 && ns != 1) || al > 13) return Fail("bad spectral selection");
			for (int i = 0; i < ns; ++i)
			{
				const bool needDC = ss == 0 && (!progressive || ah == 0), needAC = se > 0 && (!progressive || ss > 0);
				if (needDC && !dc[scan[i]->td].present) return Fail("missing DC Huffman table");
				if (needAC && !ac[scan[i]->ta].present) return Fail("missing AC Huffman table");
			}
			BitReader br{ &file[at + len], file.data() + file.size() };
			int eobRun = 0, untilRestart = restartInterval;
			auto restartIfDue = [&]() {
				if (restartInterval == 0 || --untilRestart > 0) return;
				br.Restart();
				for (int c = 0; c < numComps; ++c) comps[c].pred = 0;
				eobRun = 0; untilRestart = restartInterval; };
			for (int c = 0; c < numComps; ++c) comps[c].pred = 0;
			if (ns == 1)
			{
				Component& k = *scan[0];
				const int total = k.scanW * k.scanH;
				for (int n = 0; n < total; ++n)
				{
					const int bx = n % k.scanW, by = n / k.scanW;
					DecodeBlock(br, k, &k.coef[((size_t)by * k.blocksW + bx) * 64], ss, se, ah, al, eobRun);
					if (n + 1 < total) restartIfDue();
				}
			}
			else
			{
				const int total = mcusX * mcusY;
				for (int n = 0; n < total; ++n)
				{
					const int mx = n % mcusX, my = n / mcusX;
					for (int i = 0; i < ns; ++i)
					{
						Component& k = *scan[i];
						for (int y = 0; y < k.v; ++y)
							for (int x = 0; x < k.h; ++x)
								DecodeBlock(br, k, &k.coef[((size_t)(my * k.v + y) * k.blocksW + (mx * k.h + x)) * 64], ss, se, ah, al, eobRun);
					}
					if (n + 1 < total) restartIfDue();
				}
			}
			// the next marker: first 0xFF xx (xx not 0, not a restart marker, not fill) at or after the reader's position
			const unsigned char* p = br.p;
			const unsigned char* end = file.data() + file.size();
			while (p + 1 < end && !(p[0] == 0xFF && p[1] != 0x00 && p[1] != 0xFF && !(p[1] >= 0xD0 && p[1] <= 0xD7))) ++p;
			next = (size_t)(p - file.data());
			return true;
		}

		bool Run()
		{
			if (file.size() < 4 || file[0] != 0xFF || file[1] != 0xD8) return Fail("not a JPEG file");
			bool sawScan = false;
			while (pos + 4 <= file.size())
			{
				if (file[pos] != 0xFF) { ++pos; continue; }
				const int marker = file[pos + 1];
				if (marker == 0xFF || marker == 0x00) { ++pos; continue; }
				if (marker == 0xD9) break;
				if (marker == 0x01 || (marker >= 0xD0 && marker <= 0xD7)) { pos += 2; continue; }
				const size_t len = ((size_t)file[pos + 2] << 8) | file[pos + 3];
				if (len < 2 || pos + 2 + len > file.size()) { if (sawScan) break; return Fail("truncated segment"); }
				const size_t at = pos + 4, body = len - 2;
				if (marker == 0xC0 || marker == 0xC1 || marker == 0xC2) { if (!ParseFrame(marker, at, body)) return false; }
				else if (marker == 0xC3 || (marker >= 0xC5 && marker <= 0xCF && marker != 0xC8 && marker != 0xCC && marker != 0xC4))
					return Fail("lossless, hierarchical and arithmetic-coded JPEG are not supported");
				else if (marker == 0xDA)
				{
					size_t next = 0;
					if (!DecodeScan(at, body, next)) return false;
					sawScan = true;
					pos = next;
					continue;
				}
				else if (!ParseTables(marker, at, body)) return false;
				pos += 2 + len;
			}
			if (!sawFrame || !sawScan) return Fail("no picture data");
			return true;
		}

		void ReconstructPlanes()
		{
			for (int c = 0; c < numComps; ++c)
			{
				Component& k = comps[c];
				static const unsigned short ones[64] = { 1,1,1,1,1,1,1,1, 1,1,1,1,1,1,1,1, 1,1,1,1,1,1,1,1, 1,1,1,1,1,1,1,1, 1,1,1,1,1,1,1,1, 1,1,1,1,1,1,1,1, 1,1,1,1,1,1,1,1, 1,1,1,1,1,1,1,1 };
				const unsigned short* q = quantPresent[k.tq] ? quant[k.tq] : ones;
				const int stride = k.blocksW * 8;
				k.plane.resize((size_t)stride * k.blocksH * 8);
				for (int by = 0; by < k.blocksH; ++by)
					for (int bx = 0; bx < k.blocksW; ++bx)
						InverseDCT(&k.coef[((size_t)by * k.blocksW + bx) * 64], q, &k.plane[(size_t)by * 8 * stride + bx * 8], stride);
				std::vector<short>().swap(k.coef);
			}
		}

		// component c at full resolution (width x height), libjpeg's default upsampling
		void Upsample(int c, std::vector<unsigned char>& out) const
		{
			const Component& k = comps[c];
			const int stride = k.blocksW * 8;
			out.resize((size_t)width * height);
			const int hx = hMax / k.h, vx = vMax / k.v;
			const bool wholeH = hMax % k.h == 0, wholeV = vMax % k.v == 0;
			auto rowOf = [&](int y) { return &k.plane[(size_t)std::min(std::max(y, 0), k.sampleH - 1) * stride]; };
			if (hx == 1 && vx == 1 && wholeH && wholeV)
			{
				for (int y = 0; y < height; ++y) memcpy(&out[(size_t)y * width], rowOf(y), width);
				return;
			}
			std::vector<unsigned char> line((size_t)k.sampleW * 2 + 2);
			if (wholeH && wholeV && hx == 2 && vx == 1 && k.sampleW > 2)
			{
				// h2v1 "fancy": 3/4 nearer + 1/4 farther sample, alternating rounding
				for (int y = 0; y < height; ++y)
				{
					const unsigned char* in = rowOf(y);
					const int n = k.sampleW;
					line[0] = in[0]; line[1] = (unsigned char)((in[0] * 3 + in[1] + 2) >> 2);
					for (int x = 1; x < n - 1; ++x)
					{
						line[2 * x] = (unsigned char)((in[x] * 3 + in[x - 1] + 1) >> 2);
						line[2 * x + 1] = (unsigned char)((in[x] * 3 + in[x + 1] + 2) >> 2);
					}
					line[2 * n - 2] = (unsigned char)((in[n - 1] * 3 + in[n - 2] + 1) >> 2); line[2 * n - 1] = in[n - 1];
					memcpy(&out[(size_t)y * width], line.data(), width);
				}
				return;
			}
			if (wholeH && wholeV && hx == 2 && vx == 2 && k.sampleW > 2)
			{
				// h2v2 "fancy": 9/16, 3/16, 3/16, 1/16 of the four nearest samples; rows beyond the picture repeat the edge row
				for (int y = 0; y < height; ++y)
				{
					const int inRow = y >> 1;
					const unsigned char* in0 = rowOf(inRow);
					const unsigned char* in1 = rowOf((y & 1) ? inRow + 1 : inRow - 1);
					const int n = k.sampleW;
					int last, cur = in0[0] * 3 + in1[0], nextSum = in0[1] * 3 + in1[1];
					line[0] = (unsigned char)((cur * 4 + 8) >> 4); line[1] = (unsigned char)((cur * 3 + nextSum + 7) >> 4);
					last = cur; cur = nextSum;
					for (int x = 1; x < n - 1; ++x)
					{
						nextSum = in0[x + 1] * 3 + in1[x + 1];
						line[2 * x] = (unsigned char)((cur * 3 + last + 8) >> 4);
						line[2 * x + 1] = (unsigned char)((cur * 3 + nextSum + 7) >> 4);
						last = cur; cur = nextSum;
					}
					line[2 * n - 2] = (unsigned char)((cur * 3 + last + 8) >> 4); line[2 * n - 1] = (unsigned char)((cur * 4 + 7) >> 4);
					memcpy(&out[(size_t)y * width], line.data(), width);
				}
				return;
			}
			if (wholeH && wholeV && hx == 1 && vx == 2)
			{
				// h1v2 "fancy" (libjpeg-turbo): 3/4 nearer + 1/4 farther row
				for (int y = 0; y < height; ++y)
				{
					const int inRow = y >> 1, bias = (y & 1) ? 2 : 1;
					const unsigned char* in0 = rowOf(inRow);
					const unsigned char* in1 = rowOf((y & 1) ? inRow + 1 : inRow - 1);
					for (int x = 0; x < width; ++x) out[(size_t)y * width + x] = (unsigned char)((in0[x] * 3 + in1[x] + bias) >> 2);
				}
				return;
			}
			// everything else: sample replication (libjpeg's int_upsample; fractional ratios by nearest sample)
			for (int y = 0; y < height; ++y)
			{
				const unsigned char* in = rowOf(y * k.v / vMax);
				for (int x = 0; x < width; ++x) out[(size_t)y * width + x] = in[std::min(k.sampleW - 1, x * k.h / hMax)];
			}
		}
	};

	// ---- encoder --------------------------------------------------------------------------------------
	const unsigned char kLumaQuant[64] = {
		16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
		18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99 };
	const unsigned char kChromaQuant[64] = {
		17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
		99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99 };
	const unsigned char kDcLumaBits[16] = { 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0 };
	const unsigned char kDcChromaBits[16] = { 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0 };
	const unsigned char kDcVals[12] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 };
	const unsigned char kAcLumaBits[16] = { 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d };
	const unsigned char kAcLumaVals[162] = {
		0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
		0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
		0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
		0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
		0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
		0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
		0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa };
	const unsigned char kAcChromaBits[16] = { 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77 };
	const unsigned char kAcChromaVals[162] = {
		0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
		0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
		0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
		0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
		0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
		0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
		0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa };

	struct EncTable { unsigned short code[256]; unsigned char size[256]; };
	void MakeEncTable(const unsigned char* bits, const unsigned char* vals, EncTable& t)
	{
		memset(&t, 0, sizeof(t));
		int code = 0, k = 0;
		for (int l = 1; l <= 16; ++l)
		{
			for (int i = 0; i < bits[l - 1]; ++i, ++k, ++code) { t.code[vals[k]] = (unsigned short)code; t.size[vals[k]] = (unsigned char)l; }
			code <<= 1;
		}
	}

	struct BitWriter
	{
		Bytes& out;
		unsigned int acc = 0; int count = 0;
		void Put(unsigned int bits, int n)
		{
			acc = (acc << n) | (bits & ((1u << n) - 1u)); count += n;
			while (count >= 8)
			{
				const unsigned char b = (unsigned char)(acc >> (count - 8));
				out.push_back(b);
				if (b == 0xFF) out.push_back(0);
				count -= 8;
			}
		}
		void Flush() { if (count) Put(0x7F, 8 - count); }
	};

	void ForwardDCT(const float* in, float* out)
	{
		static float basis[8][8];
		static bool ready = false;
		if (!ready)
		{
			for (int u = 0; u < 8; ++u)
				for (int x = 0; x < 8; ++x)
					basis[u][x] = (float)((u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0));
			ready = true;
		}
		float tmp[64];
		for (int y = 0; y < 8; ++y)
			for (int u = 0; u < 8; ++u)
			{
				float s = 0.0f;
				for (int x = 0; x < 8; ++x) s += in[8 * y + x] * basis[u][x];
				tmp[8 * y + u] = s;
			}
		for (int v = 0; v < 8; ++v)
			for (int u = 0; u < 8; ++u)
			{
				float s = 0.0f;
				for (int y = 0; y < 8; ++y) s += tmp[8 * y + u] * basis[v][y];
				out[8 * v + u] = s;
			}
	}

	void EncodeBlock(BitWriter& bw, const float* samples, const unsigned char* quant, int& pred, const EncTable& dcT, const EncTable& acT)
	{
		float freq[64];
		ForwardDCT(samples, freq);
		int q[64];
		for (int i = 0; i < 64; ++i) q[i] = (int)std::lround(freq[kZigzag[i]] / (float)quant[kZigzag[i]]);
		auto category = [](int v) { int a = v < 0 ? -v : v, s = 0; while (a) { ++s; a >>= 1; } return s; };
		const int diff = q[0] - pred; pred = q[0];
		int s = category(diff);
		bw.Put(dcT.code[s], dcT.size[s]);
		if (s) bw.Put((unsigned int)(diff < 0 ? diff - 1 : diff), s);
		int run = 0;
		for (int i = 1; i < 64; ++i)
		{
			if (q[i] == 0) { ++run; continue; }
			while (run > 15) { bw.Put(acT.code[0xF0], acT.size[0xF0]); run -= 16; }
			s = category(q[i]);
			bw.Put(acT.code[(run << 4) | s], acT.size[(run << 4) | s]);
			bw.Put((unsigned int)(q[i] < 0 ? q[i] - 1 : q[i]), s);
			run = 0;
		}
		if (run) bw.Put(acT.code[0], acT.size[0]);
	}
}

static Image2D* FinishJPEG(Decoder& d);

// Image2D from a JPEG file image; nullptr (and the reason in `why`) when this is not a JPEG the decoder handles.
Image2D* RtLoadJPEG(const std::vector<unsigned char>& file, const char** why)
{
	Decoder d(file);
	try
	{
		if (!d.Run()) { if (why) *why = d.error ? d.error : "not a JPEG file"; return nullptr; }
		return FinishJPEG(d);
	}
	catch (const std::bad_alloc&) { if (why) *why = "out of memory (header announces a picture too large to hold)"; return nullptr; }
}

static Image2D* FinishJPEG(Decoder& d)
{
	d.ReconstructPlanes();
	const int w = d.width, h = d.height;
	std::vector<unsigned char> plane[3];
	for (int c = 0; c < d.numComps; ++c) d.Upsample(c, plane[c]);
	// three components are YCbCr unless an Adobe marker says "no transform" or the component ids spell RGB (libjpeg's rules)
	bool ycc = d.numComps == 3;
	if (ycc)
	{
		if (d.jfif) ycc = true;
		else if (d.adobeTransform >= 0) ycc = d.adobeTransform != 0;
		else if (d.comps[0].id == 'R' && d.comps[1].id == 'G' && d.comps[2].id == 'B') ycc = false;
	}
	Image2D* image = new Image2D((uint32)w, (uint32)h);
	Pixel* dst = image->MutablePixels();
	auto clamp8 = [](int v) { return (uint8)std::min(255, std::max(0, v)); };
	for (size_t i = 0; i < (size_t)w * h; ++i)
	{
		if (d.numComps == 1) { const uint8 y = plane[0][i]; dst[i] = Pixel(y, y, y, (uint8)255); continue; }
		if (!ycc) { dst[i] = Pixel((uint8)plane[0][i], (uint8)plane[1][i], (uint8)plane[2][i], (uint8)255); continue; }
		// jdcolor.c: 16-bit fixed point, the two green terms share one rounding
		const int y = plane[0][i], cb = (int)plane[1][i] - 128, cr = (int)plane[2][i] - 128;
		const int r = y + ((91881 * cr + 32768) >> 16);
		const int g = y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
		const int b = y + ((116130 * cb + 32768) >> 16);
		dst[i] = Pixel(clamp8(r), clamp8(g), clamp8(b), (uint8)255);
	}
	return image;
}

// Baseline JFIF file, 4:2:0, quality 1..100 (75 = what FreeImage::Save(..., 0) writes for the reference)
bool RtWriteJPEG(const Image2D* image, const char* path, int quality)
{
	const int w = (int)image->GetWidth(), h = (int)image->GetHeight();
	if (w <= 0 || h <= 0 || w > 65535 || h > 65535) return false;
	quality = std::min(100, std::max(1, quality));
	const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
	unsigned char quant[2][64];
	for (int i = 0; i < 64; ++i)
	{
		quant[0][i] = (unsigned char)std::min(255, std::max(1, (kLumaQuant[i] * scale + 50) / 100));
		quant[1][i] = (unsigned char)std::min(255, std::max(1, (kChromaQuant[i] * scale + 50) / 100));
	}
	// planes padded to whole 16x16 MCUs by repeating the edge
	const int mcusX = (w + 15) / 16, mcusY = (h + 15) / 16, pw = mcusX * 16, ph = mcusY * 16;
	std::vector<float> Y((size_t)pw * ph), Cb((size_t)pw * ph / 4), Cr((size_t)pw * ph / 4);
	{
		std::vector<float> cbFull((size_t)pw * ph), crFull((size_t)pw * ph);
		for (int y = 0; y < ph; ++y)
			for (int x = 0; x < pw; ++x)
			{
				const uint32 argb = image->GetPixel(std::min(x, w - 1), std::min(y, h - 1)).ToUint32();
				const float r = (float)((argb >> 16) & 0xFF), g = (float)((argb >> 8) & 0xFF), b = (float)(argb & 0xFF);
				const size_t i = (size_t)y * pw + x;
				Y[i] = 0.299f * r + 0.587f * g + 0.114f * b - 128.0f;
				cbFull[i] = -0.168735892f * r - 0.331264108f * g + 0.5f * b;
				crFull[i] = 0.5f * r - 0.418687589f * g - 0.081312411f * b;
			}
		for (int y = 0; y < ph / 2; ++y)
			for (int x = 0; x < pw / 2; ++x)
			{
				const size_t a = (size_t)(2 * y) * pw + 2 * x, b = a + pw;
				Cb[(size_t)y * (pw / 2) + x] = 0.25f * (cbFull[a] + cbFull[a + 1] + cbFull[b] + cbFull[b + 1]);
				Cr[(size_t)y * (pw / 2) + x] = 0.25f * (crFull[a] + crFull[a + 1] + crFull[b] + crFull[b + 1]);
			}
	}
	Bytes out;
	auto put16 = [&](int v) { out.push_back((unsigned char)(v >> 8)); out.push_back((unsigned char)v); };
	auto marker = [&](int m, int len) { out.push_back(0xFF); out.push_back((unsigned char)m); put16(len); };
	out.push_back(0xFF); out.push_back(0xD8);
	marker(0xE0, 16);
	{ const unsigned char jfif[14] = { 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0 }; out.insert(out.end(), jfif, jfif + 14); }
	for (int t = 0; t < 2; ++t)
	{
		marker(0xDB, 67);
		out.push_back((unsigned char)t);
		for (int i = 0; i < 64; ++i) out.push_back(quant[t][kZigzag[i]]);
	}
	marker(0xC0, 17);
	out.push_back(8); put16(h); put16(w); out.push_back(3);
	{ const unsigned char comps[9] = { 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1 }; out.insert(out.end(), comps, comps + 9); }
	auto huffman = [&](int id, const unsigned char* bits, const unsigned char* vals, int count) {
		marker(0xC4, 19 + count);
		out.push_back((unsigned char)id);
		out.insert(out.end(), bits, bits + 16);
		out.insert(out.end(), vals, vals + count); };
	huffman(0x00, kDcLumaBits, kDcVals, 12); huffman(0x10, kAcLumaBits, kAcLumaVals, 162);
	huffman(0x01, kDcChromaBits, kDcVals, 12); huffman(0x11, kAcChromaBits, kAcChromaVals, 162);
	marker(0xDA, 12);
	{ const unsigned char scan[10] = { 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0 }; out.insert(out.end(), scan, scan + 10); }

	EncTable dcLuma, acLuma, dcChroma, acChroma;
	MakeEncTable(kDcLumaBits, kDcVals, dcLuma); MakeEncTable(kAcLumaBits, kAcLumaVals, acLuma);
	MakeEncTable(kDcChromaBits, kDcVals, dcChroma); MakeEncTable(kAcChromaBits, kAcChromaVals, acChroma);
	BitWriter bw{ out };
	int predY = 0, predCb = 0, predCr = 0;
	float block[64];
	for (int my = 0; my < mcusY; ++my)
		for (int mx = 0; mx < mcusX; ++mx)
		{
			for (int k = 0; k < 4; ++k)
			{
				const int x0 = mx * 16 + (k & 1) * 8, y0 = my * 16 + (k >> 1) * 8;
				for (int y = 0; y < 8; ++y) memcpy(block + 8 * y, &Y[(size_t)(y0 + y) * pw + x0], 8 * sizeof(float));
				EncodeBlock(bw, block, quant[0], predY, dcLuma, acLuma);
			}
			for (int y = 0; y < 8; ++y) memcpy(block + 8 * y, &Cb[(size_t)(my * 8 + y) * (pw / 2) + mx * 8], 8 * sizeof(float));
			EncodeBlock(bw, block, quant[1], predCb, dcChroma, acChroma);
			for (int y = 0; y < 8; ++y) memcpy(block + 8 * y, &Cr[(size_t)(my * 8 + y) * (pw / 2) + mx * 8], 8 * sizeof(float));
			EncodeBlock(bw, block, quant[1], predCr, dcChroma, acChroma);
		}
	bw.Flush();
	out.push_back(0xFF); out.push_back(0xD9);
	FILE* f = fopen(path, "wb");
	if (!f) return false;
	const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
	return fclose(f) == 0 && ok;
}

// BenchRunTable: the per-frame statistics table AutoBench writes (AutoBench/benchruntable.{h,cpp}): one worksheet per
// scene, 18 columns (benchruntable.h:28-49), one row per frame, saved as <output>/Stats.xlsx.
//
// The reference goes through OpenXLSX (FetchContent, not available here). An .xlsx file is a ZIP archive of a few small
// XML parts; this writer emits exactly those parts with the "stored" (uncompressed) ZIP method, so it needs no zlib:
// [Content_Types].xml, _rels/.rels, xl/workbook.xml, xl/_rels/workbook.xml.rels and one xl/worksheets/sheetN.xml per
// worksheet, header cells as inline strings, values as numbers. Like OpenXLSX's XLDocument::create, the workbook starts
// with an empty "Sheet1" that the reference never removes (benchruntable.cpp:15-41); scenes follow in std::map order.
#ifndef BENCHRUNTABLE_H
#define BENCHRUNTABLE_H

#include <cstdint>
#include <cstdio>
#include <filesystem>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "flipsolver2d.h"

class BenchRunTable
{
public:
    enum TableColumn : int
    {
        STEP_NUMBER_COLUMN = 0,
        SUBSTEP_COUNT_COLUMN,
        TOTAL_FRAME_TIME_COLUMN,
        ADVECTION_TIME_COLUMN,
        DECOMPOSITION_TIME_COLUMN,
        DENSITY_TIME_COLUMN,
        PARTICLE_REBIN_TIME_COLUMN,
        PARTICLE_TO_GRID_TIME_COLUMN,
        GRID_UPDATE_TIME_COLUMN,
        AFTER_TRANSFER_TIME_COLUMN,
        PRESSURE_TIME_COLUMN,
        VISCOSITY_TIME_COLUMN,
        REPRESSURE_TIME_COLUMN,
        PARTICLE_UPDATE_TIME_COLUMN,
        PARTICLE_RESEED_TIME_COLUMN,
        PRESSURE_ITERS_COLUMN,
        DENSITY_ITERS_COLUMN,
        VISCOSITY_ITERS_COLUMN,
        TABLE_COLUMN_COUNT
    };

    // one table row as numbers (benchruntable.cpp:64-84); step = 1-based frame number
    using Row = std::array<double, TABLE_COLUMN_COUNT>;

    void setOutputFile(const std::filesystem::path &p) { m_outputFilePath = p; }
    void addStepTiming(const SolverStats &stats) { m_currentSceneRows.push_back(rowOf(stats, static_cast<int>(m_currentSceneRows.size()) + 1)); }
    void addRow(const Row &r) { m_currentSceneRows.push_back(r); }
    void finishScene(const std::string &sceneName)
    {
        m_allSceneRows.insert({sceneName, m_currentSceneRows});
        m_currentSceneRows.clear();
    }
    const std::map<std::string, std::vector<Row>> &scenes() const { return m_allSceneRows; }

    static Row rowOf(const SolverStats &s, int step)
    {
        const SolverStats::StageTimings t = s.timings();
        Row r{};
        r[STEP_NUMBER_COLUMN] = step;
        r[SUBSTEP_COUNT_COLUMN] = s.substepCount();
        r[TOTAL_FRAME_TIME_COLUMN] = s.frameTime();
        r[ADVECTION_TIME_COLUMN] = t[ADVECTION];
        r[DECOMPOSITION_TIME_COLUMN] = t[DECOMPOSITION];
        r[DENSITY_TIME_COLUMN] = t[DENSITY];
        r[PARTICLE_REBIN_TIME_COLUMN] = t[PARTICLE_REBIN];
        r[PARTICLE_TO_GRID_TIME_COLUMN] = t[PARTICLE_TO_GRID];
        r[GRID_UPDATE_TIME_COLUMN] = t[GRID_UPDATE];
        r[AFTER_TRANSFER_TIME_COLUMN] = t[AFTER_TRANSFER];
        r[PRESSURE_TIME_COLUMN] = t[PRESSURE];
        r[VISCOSITY_TIME_COLUMN] = t[VISCOSITY];
        r[REPRESSURE_TIME_COLUMN] = t[REPRESSURE];
        r[PARTICLE_UPDATE_TIME_COLUMN] = t[PARTICLE_UPDATE];
        r[PARTICLE_RESEED_TIME_COLUMN] = t[PARTICLE_RESEED];
        r[PRESSURE_ITERS_COLUMN] = s.pressureIterations();
        r[DENSITY_ITERS_COLUMN] = s.densityIterations();
        r[VISCOSITY_ITERS_COLUMN] = s.viscosityIterations();
        return r;
    }

    // benchruntable.cpp:87-145
    static const char *getColumnHeader(int column)
    {
        static const char *names[TABLE_COLUMN_COUNT] = {
            "Step number",     "Substeps",         "Frame time",       "Advection",           "Decomposition",      "Density correction",
            "Particle rebin",  "Particle to grid", "Grid update",      "After transfer",      "Pressure",           "Viscosity",
            "After-visc pressure", "Particle update", "Particle reseeding", "Pressure iterations", "Density iterations", "Viscosity iterations"};
        return column >= 0 && column < TABLE_COLUMN_COUNT ? names[column] : "INVALID COLUMN";
    }

    // Stats.xlsx; returns false when the file cannot be written
    bool save() const
    {
        std::vector<std::pair<std::string, std::string>> parts;  // (path inside the archive, content)
        std::vector<std::string> sheetNames = {"Sheet1"};
        for (const auto &it : m_allSceneRows) sheetNames.push_back(sheetName(it.first, sheetNames));
        std::ostringstream types, wb, wbRels;
        types << "<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"yes\"?>\n<Types xmlns=\"http://schemas.openxmlformats.org/package/2006/content-types\">"
              << "<Default Extension=\"rels\" ContentType=\"application/vnd.openxmlformats-package.relationships+xml\"/>"
              << "<Default Extension=\"xml\" ContentType=\"application/xml\"/>"
              << "<Override PartName=\"/xl/workbook.xml\" ContentType=\"application/vnd.openxmlformats-officedocument.spreadsheetml.sheet.main+xml\"/>";
        wb << "<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"yes\"?>\n<workbook xmlns=\"http://schemas.openxmlformats.org/spreadsheetml/2006/main\" "
           << "xmlns:r=\"http://schemas.openxmlformats.org/officeDocument/2006/relationships\"><sheets>";
        wbRels << "<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"yes\"?>\n<Relationships xmlns=\"http://schemas.openxmlformats.org/package/2006/relationships\">";
        for (size_t k = 0; k < sheetNames.size(); k++)
        {
            const std::string id = std::to_string(k + 1);
            types << "<Override PartName=\"/xl/worksheets/sheet" << id
                  << ".xml\" ContentType=\"application/vnd.openxmlformats-officedocument.spreadsheetml.worksheet+xml\"/>";
            wb << "<sheet name=\"" << escape(sheetNames[k]) << "\" sheetId=\"" << id << "\" r:id=\"rId" << id << "\"/>";
            wbRels << "<Relationship Id=\"rId" << id
                   << "\" Type=\"http://schemas.openxmlformats.org/officeDocument/2006/relationships/worksheet\" Target=\"worksheets/sheet" << id << ".xml\"/>";
        }
        types << "</Types>";
        wb << "</sheets></workbook>";
        wbRels << "</Relationships>";
        parts.push_back({"[Content_Types].xml", types.str()});
        parts.push_back({"_rels/.rels",
                         "<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"yes\"?>\n<Relationships "
                         "xmlns=\"http://schemas.openxmlformats.org/package/2006/relationships\"><Relationship Id=\"rId1\" "
                         "Type=\"http://schemas.openxmlformats.org/officeDocument/2006/relationships/officeDocument\" "
                         "Target=\"xl/workbook.xml\"/></Relationships>"});
        parts.push_back({"xl/workbook.xml", wb.str()});
        parts.push_back({"xl/_rels/workbook.xml.rels", wbRels.str()});
        parts.push_back({"xl/worksheets/sheet1.xml", sheetXml(nullptr)});
        size_t k = 2;
        for (const auto &it : m_allSceneRows) parts.push_back({"xl/worksheets/sheet" + std::to_string(k++) + ".xml", sheetXml(&it.second)});
        return writeZip(m_outputFilePath, parts);
    }

private:
    static std::string escape(const std::string &s)
    {
        std::string o;
        for (char c : s)
        {
            if (c == '&') o += "&amp;";
            else if (c == '<') o += "&lt;";
            else if (c == '>') o += "&gt;";
            else if (c == '"') o += "&quot;";
            else o += c;
        }
        return o;
    }

    // worksheet names: at most 31 characters, none of []:*?/\ and unique
    static std::string sheetName(const std::string &scene, const std::vector<std::string> &taken)
    {
        std::string n;
        for (char c : scene) n += (c == '[' || c == ']' || c == ':' || c == '*' || c == '?' || c == '/' || c == '\\') ? '_' : c;
        if (n.empty()) n = "scene";
        if (n.size() > 31) n.resize(31);
        std::string candidate = n;
        for (int suffix = 2;; suffix++)
        {
            bool clash = false;
            for (const std::string &t : taken) clash = clash || t == candidate;
            if (!clash) return candidate;
            const std::string tail = "_" + std::to_string(suffix);
            candidate = n.substr(0, 31 - tail.size()) + tail;
        }
    }

    static std::string columnLetters(int c)  // 0 -> A, 25 -> Z, 26 -> AA
    {
        std::string s;
        for (c += 1; c > 0; c = (c - 1) / 26) s.insert(s.begin(), static_cast<char>('A' + (c - 1) % 26));
        return s;
    }

    static std::string sheetXml(const std::vector<Row> *rows)
    {
        std::ostringstream x;
        x.precision(9);
        x << "<?xml version=\"1.0\" encoding=\"UTF-8\" standalone=\"yes\"?>\n<worksheet xmlns=\"http://schemas.openxmlformats.org/spreadsheetml/2006/main\"><sheetData>";
        if (rows)
        {
            x << "<row r=\"1\">";
            for (int c = 0; c < TABLE_COLUMN_COUNT; c++)
                x << "<c r=\"" << columnLetters(c) << "1\" t=\"inlineStr\"><is><t>" << escape(getColumnHeader(c)) << "</t></is></c>";
            x << "</row>";
            int r = 2;
            for (const Row &row : *rows)
            {
                x << "<row r=\"" << r << "\">";
                for (int c = 0; c < TABLE_COLUMN_COUNT; c++) x << "<c r=\"" << columnLetters(c) << r << "\"><v>" << row[static_cast<size_t>(c)] << "</v></c>";
                x << "</row>";
                r++;
            }
        }
        x << "</sheetData></worksheet>";
        return x.str();
    }

    static uint32_t crc32(const std::string &data)
    {
        static uint32_t table[256];
        static bool ready = false;
        if (!ready)
        {
            for (uint32_t n = 0; n < 256; n++)
            {
                uint32_t c = n;
                for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
                table[n] = c;
            }
            ready = true;
        }
        uint32_t c = 0xFFFFFFFFu;
        for (unsigned char b : data) c = table[(c ^ b) & 0xFFu] ^ (c >> 8);
        return c ^ 0xFFFFFFFFu;
    }

    static void put16(std::string &o, uint32_t v)
    {
        o.push_back(static_cast<char>(v & 0xFF));
        o.push_back(static_cast<char>((v >> 8) & 0xFF));
    }
    static void put32(std::string &o, uint32_t v)
    {
        put16(o, v & 0xFFFF);
        put16(o, v >> 16);
    }

    // ZIP with method 0 (stored): local headers + data, central directory, end record (APPNOTE 4.3)
    static bool writeZip(const std::filesystem::path &path, const std::vector<std::pair<std::string, std::string>> &parts)
    {
        std::string out, central;
        for (const auto &p : parts)
        {
            const uint32_t crc = crc32(p.second), size = static_cast<uint32_t>(p.second.size()), offset = static_cast<uint32_t>(out.size());
            put32(out, 0x04034b50u);
            put16(out, 20);  // version needed
            put16(out, 0);   // flags
            put16(out, 0);   // method: stored
            put16(out, 0);   // time
            put16(out, 0x21);  // date: 1980-01-01
            put32(out, crc);
            put32(out, size);
            put32(out, size);
            put16(out, static_cast<uint32_t>(p.first.size()));
            put16(out, 0);
            out += p.first;
            out += p.second;
            put32(central, 0x02014b50u);
            put16(central, 20);
            put16(central, 20);
            put16(central, 0);
            put16(central, 0);
            put16(central, 0);
            put16(central, 0x21);
            put32(central, crc);
            put32(central, size);
            put32(central, size);
            put16(central, static_cast<uint32_t>(p.first.size()));
            put16(central, 0);
            put16(central, 0);
            put16(central, 0);
            put16(central, 0);
            put32(central, 0);
            put32(central, offset);
            central += p.first;
        }
        const uint32_t cdOffset = static_cast<uint32_t>(out.size()), cdSize = static_cast<uint32_t>(central.size());
        out += central;
        put32(out, 0x06054b50u);
        put16(out, 0);
        put16(out, 0);
        put16(out, static_cast<uint32_t>(parts.size()));
        put16(out, static_cast<uint32_t>(parts.size()));
        put32(out, cdSize);
        put32(out, cdOffset);
        put16(out, 0);
        FILE *f = std::fopen(path.string().c_str(), "wb");
        if (!f) return false;
        const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
        std::fclose(f);
        return ok;
    }

    std::filesystem::path m_outputFilePath = "Stats.xlsx";
    std::vector<Row> m_currentSceneRows;
    std::map<std::string, std::vector<Row>> m_allSceneRows;
};

#endif
